// Runtime shim: the same kernel sources build (a) with nvcc for sm_100a -- the product -- and
// (b) with g++ as a sequential/multi-threaded host emulation used ONLY by the CPU test-suite
// (tests/_emul/, -DZKFL_EMUL) to check kernel logic where no GPU exists.  The emulation is never
// built into, loaded by, or reachable from the shipped library: libzkfl.so has no CPU path.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifndef ZKFL_EMUL
// ------------------------------------------------------------------ CUDA build
#include <cuda_runtime.h>
#define ZK_HD __host__ __device__ __forceinline__
#define ZK_D __device__ __forceinline__
#define ZK_HD_NOINLINE __host__ __device__ __noinline__
// kernels have internal linkage: kernels.cuh is included by several .cu files and each compiles only the kernels it launches
#define ZK_GLOBAL static __global__
#define ZK_UNROLL _Pragma("unroll")
#define ZK_NOUNROLL _Pragma("unroll 1")
#define ZK_TID ((size_t)blockIdx.x * blockDim.x + threadIdx.x)
#define ZK_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define ZK_ATOMIC_OR(p, v) atomicOr((p), (v))
#define ZK_LDG(p) __ldg(p)

namespace zkrt {
typedef cudaStream_t stream_t;
inline const char* err_str(cudaError_t e) { return cudaGetErrorString(e); }
}  // namespace zkrt

// launches `kernel` over at least `total` threads, `block` threads per CTA, on `stream`
#define ZK_LAUNCH(kernel, total, block, stream, ...)                                         \
  do {                                                                                       \
    size_t _tot = (size_t)(total);                                                           \
    if (_tot) {                                                                              \
      unsigned _grid = (unsigned)((_tot + (block)-1) / (block));                             \
      kernel<<<_grid, (block), 0, (stream)>>>(__VA_ARGS__);                                  \
      zkrt::note_launch(#kernel);                                                            \
      if (zkrt::debug_sync()) zkrt::debug_check(#kernel, (stream));                          \
    }                                                                                        \
  } while (0)

#else
// ------------------------------------------------------------------ host emulation (tests only)
#include <atomic>
#include <cstdlib>
#include <thread>
#include <vector>
#define ZK_HD inline
#define ZK_D inline
#define ZK_HD_NOINLINE inline
#define ZK_GLOBAL static
#define ZK_UNROLL
#define ZK_NOUNROLL
#define __restrict__
struct zk_emul_idx { size_t tid; };
extern thread_local zk_emul_idx zk_emul_cur;
#define ZK_TID (zk_emul_cur.tid)
#define ZK_ATOMIC_ADD(p, v) __atomic_fetch_add((p), (v), __ATOMIC_RELAXED)
#define ZK_ATOMIC_OR(p, v) __atomic_fetch_or((p), (v), __ATOMIC_RELAXED)
#define ZK_LDG(p) (*(p))

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
#define cudaSuccess 0
enum { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
namespace zkrt {
typedef void* stream_t;
inline const char* err_str(int) { return "emul"; }
}  // namespace zkrt
inline int cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
inline int cudaFree(void* p) { free(p); return 0; }
inline int cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
inline int cudaFreeHost(void* p) { free(p); return 0; }
inline int cudaMemcpyAsync(void* d, const void* s, size_t n, int, void*) { memcpy(d, s, n); return 0; }
inline int cudaMemsetAsync(void* d, int v, size_t n, void*) { memset(d, v, n); return 0; }
inline int cudaStreamSynchronize(void*) { return 0; }
inline int cudaStreamCreate(void** s) { *s = nullptr; return 0; }
inline int cudaStreamDestroy(void*) { return 0; }
inline int cudaSetDevice(int) { return 0; }
inline int cudaGetLastError() { return 0; }
inline int cudaEventCreate(void** e) { *e = nullptr; return 0; }
inline int cudaEventDestroy(void*) { return 0; }
inline int cudaEventRecord(void*, void*) { return 0; }
inline int cudaStreamWaitEvent(void*, void*, unsigned) { return 0; }
inline int cudaEventSynchronize(void*) { return 0; }
inline int cudaEventElapsedTime(float* ms, void*, void*) { *ms = 0.f; return 0; }
inline int cudaGetDeviceCount(int* n) { *n = 1; return 0; }

#define ZK_LAUNCH(kernel, total, block, stream, ...)                                          \
  do {                                                                                        \
    size_t _tot = (size_t)(total);                                                            \
    if (_tot) {                                                                               \
      size_t _padded = (_tot + (block)-1) / (block) * (block);                                \
      zkrt::emul_run(_padded, [&](size_t _i) { zk_emul_cur.tid = _i; kernel(__VA_ARGS__); }); \
      zkrt::note_launch(#kernel);                                                             \
    }                                                                                         \
  } while (0)

namespace zkrt {
template <class Fn>
inline void emul_run(size_t total, Fn fn) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  if (nt > 16) nt = 16;
  if (total < 4096) nt = 1;
  if (nt == 1) { for (size_t i = 0; i < total; i++) fn(i); return; }
  std::atomic<size_t> next(0);
  const size_t chunk = 256;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([&] {
      for (;;) {
        size_t s = next.fetch_add(chunk);
        if (s >= total) break;
        size_t e = s + chunk < total ? s + chunk : total;
        for (size_t i = s; i < e; i++) fn(i);
      }
    });
  for (auto& t : th) t.join();
}
}  // namespace zkrt
#endif

namespace zkrt {
void note_launch(const char* name);  // defined in zkfl.cu: counts launches for gpu_launches / profiling
bool debug_sync();                   // ZKFL_DEBUG_SYNC=1: synchronise and check after every launch
void debug_check(const char* name, cudaStream_t stream);
}
