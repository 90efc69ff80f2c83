# dev: the host emulation of the kernels (tests/_emul build flags + AddressSanitizer + UBSan), main paths of the library.
# Catches out-of-bounds indexing in kernel logic that a passing GPU run would not reveal.  Objects under build/asan/ (git-ignored).
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/asan
CS=verifiable-federated-training-with-zero-knowledge-proofs-zk-fl-_b200/csrc
for u in zkfl witness msm_g1 msm_g2 verify setup; do
  g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -DZKFL_EMUL -x c++ -std=c++17 -fPIC -pthread \
      -Wno-unknown-pragmas -Wno-unused-function -c -o build/asan/$u.o $CS/$u.cu &
done
wait
g++ -shared -pthread -fsanitize=address,undefined -o build/asan/libzkfl_asan.so build/asan/*.o
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) python tests/dev/asan_emul.py
