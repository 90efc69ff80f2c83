{
  "targets": [{
    "target_name": "zkfl_napi",
    "sources": ["zkfl_napi.cc"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/../../verifiable-federated-training-with-zero-knowledge-proofs-zk-fl-_b200", "-lzkfl",
                  "-Wl,-rpath,<(module_root_dir)/../../verifiable-federated-training-with-zero-knowledge-proofs-zk-fl-_b200"],
    "cflags_cc": ["-std=c++17", "-O2"],
    "defines": ["NAPI_VERSION=8"]
  }]
}
