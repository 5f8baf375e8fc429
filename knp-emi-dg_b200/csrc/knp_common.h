// knp_common.h - launch / memory layer shared by all translation units.
//
// Two builds exist from the same sources:
//   * nvcc -gencode arch=compute_100a,code=sm_100a  -> libknpemi.so, the product;
//   * g++ -DKNP_EMU                                  -> tests/emu/libknpemi_emu.so, a
//     host emulation that runs the per-thread kernel bodies in a loop.  It exists
//     only so that the CPU test-suite (`pytest -m "not gpu"`) can check the host
//     logic and the kernel arithmetic against the oracle without a GPU.  The Python
//     package never loads it (knpemidg/_lib.py only looks for libknpemi.so and
//     raises if it is missing).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <stdexcept>
#include <atomic>

#ifdef KNP_EMU
#define KNP_HD inline
#define KNP_D inline
typedef int knp_stream_t;
#else
#include <cuda_runtime.h>
#define KNP_HD __host__ __device__ __forceinline__
#define KNP_D __device__ __forceinline__
typedef cudaStream_t knp_stream_t;
#endif

namespace knp {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

[[noreturn]] inline void fail(const std::string& msg) { throw Error(msg); }

#ifndef KNP_EMU
inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %s at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    throw Error(buf);
  }
}
#define KNP_CUDA(x) ::knp::cuda_check((x), #x, __FILE__, __LINE__)
#endif

// ---- memory ---------------------------------------------------------------
inline void* dev_alloc_bytes(size_t bytes) {
  if (bytes == 0) bytes = 8;
#ifdef KNP_EMU
  void* p = calloc(1, bytes);
  if (!p) fail("host emulation: out of memory");
  return p;
#else
  void* p = nullptr;
  KNP_CUDA(cudaMalloc(&p, bytes));
  // the memset runs on the legacy default stream, the library's streams are non-blocking:
  // wait for it, or it may land after the first kernel that writes the buffer
  KNP_CUDA(cudaMemset(p, 0, bytes));
  KNP_CUDA(cudaStreamSynchronize(0));
  return p;
#endif
}
inline void dev_free(void* p) {
  if (!p) return;
#ifdef KNP_EMU
  free(p);
#else
  cudaFree(p);
#endif
}
inline void h2d(void* dst, const void* src, size_t bytes, knp_stream_t s) {
  if (!bytes) return;
#ifdef KNP_EMU
  (void)s;
  memcpy(dst, src, bytes);
#else
  KNP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
  KNP_CUDA(cudaStreamSynchronize(s));
#endif
}
inline void d2h(void* dst, const void* src, size_t bytes, knp_stream_t s) {
  if (!bytes) return;
#ifdef KNP_EMU
  (void)s;
  memcpy(dst, src, bytes);
#else
  KNP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
  KNP_CUDA(cudaStreamSynchronize(s));
#endif
}
inline void d2d(void* dst, const void* src, size_t bytes, knp_stream_t s) {
  if (!bytes) return;
#ifdef KNP_EMU
  (void)s;
  memmove(dst, src, bytes);
#else
  KNP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
#endif
}
inline void dev_zero(void* dst, size_t bytes, knp_stream_t s) {
  if (!bytes) return;
#ifdef KNP_EMU
  (void)s;
  memset(dst, 0, bytes);
#else
  KNP_CUDA(cudaMemsetAsync(dst, 0, bytes, s));
#endif
}
inline void stream_sync(knp_stream_t s) {
#ifdef KNP_EMU
  (void)s;
#else
  KNP_CUDA(cudaStreamSynchronize(s));
#endif
}

// Owning device array.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { dev_free(p); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { dev_free(p); }
  void alloc(size_t count) {
    dev_free(p);
    p = static_cast<T*>(dev_alloc_bytes(count * sizeof(T)));
    n = count;
  }
  void upload(const T* src, size_t count, knp_stream_t s) {
    if (count != n) alloc(count);
    h2d(p, src, count * sizeof(T), s);
  }
  void upload(const std::vector<T>& v, knp_stream_t s) { upload(v.data(), v.size(), s); }
  std::vector<T> download(knp_stream_t s) const {
    std::vector<T> v(n);
    d2h(v.data(), p, n * sizeof(T), s);
    return v;
  }
};

// number of kernel launches issued by this library (bench.py reports it as gpu_launches)
inline std::atomic<long long>& launch_counter() {
  static std::atomic<long long> n{0};
  return n;
}

// ---- launch ---------------------------------------------------------------
// One-thread-per-index kernels are functor structs with `void operator()(int64_t) const`;
// the functor type names the kernel in profiles (knp::pf_kernel<knp::EmiCellKernel<3>>).
#ifdef KNP_EMU
// (-fopenmp: the C++/OpenMP port that bench.py times as the CPU baseline; the test-suite's
// emulation library is built without it and stays serial and bit-reproducible)
template <class F>
inline void parallel_for(knp_stream_t, int64_t n, const F& f, int = 256) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (n > 2048)
#endif
  for (int64_t i = 0; i < n; ++i) f(i);
}
#else
template <class F>
__global__ void pf_kernel(int64_t n, const F f) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) f(i);
}
template <class F>
inline void parallel_for(knp_stream_t s, int64_t n, const F& f, int block = 256) {
  if (n <= 0) return;
  int64_t grid = (n + block - 1) / block;
  ++launch_counter();
  pf_kernel<F><<<(unsigned)grid, block, 0, s>>>(n, f);
  KNP_CUDA(cudaGetLastError());
}
#endif

// ---- batched launch ---------------------------------------------------------
// The same one-thread-per-index functor for up to MAX_BATCH independent data sets (the solved
// ions' linear systems: same mesh, same sparsity, different matrices and vectors) in ONE launch:
// blockIdx.y selects the functor.  Halves (N_ions = 2) the launch count and the latency-bound
// tails of the KNP solve, which treats the ions as one block-diagonal system like the
// reference's mixed space (solver.py:168-169).
constexpr int MAX_BATCH = 6;
template <class F>
struct BatchOf { F f[MAX_BATCH]; };
#ifdef KNP_EMU
template <class F>
inline void parallel_for_batch(knp_stream_t s, int64_t n, int nb, const BatchOf<F>& b, int block = 256) {
  for (int k = 0; k < nb; ++k) parallel_for(s, n, b.f[k], block);
}
template <class F>
inline void parallel_for_batch_range(knp_stream_t, int64_t lo, int64_t hi, int nb, const BatchOf<F>& b, int = 256) {
  for (int k = 0; k < nb; ++k)
    for (int64_t i = lo; i < hi; ++i) b.f[k](i);
}
#else
template <class F>
__global__ void pf_batch_kernel(int64_t n, const BatchOf<F> b) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) b.f[blockIdx.y](i);
}
// indices lo .. hi-1 only (lo a multiple of the block size's lane grouping: callers pass multiples of ND)
template <class F>
__global__ void pf_batch_range_kernel(int64_t lo, int64_t hi, const BatchOf<F> b) {
  int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < hi) b.f[blockIdx.y](i);
}
template <class F>
inline void parallel_for_batch_range(knp_stream_t s, int64_t lo, int64_t hi, int nb, const BatchOf<F>& b, int block = 256) {
  if (hi <= lo || nb <= 0) return;
  const int64_t grid = (hi - lo + block - 1) / block;
  ++launch_counter();
  pf_batch_range_kernel<F><<<dim3((unsigned)grid, (unsigned)nb), block, 0, s>>>(lo, hi, b);
  KNP_CUDA(cudaGetLastError());
}
template <class F>
inline void parallel_for_batch(knp_stream_t s, int64_t n, int nb, const BatchOf<F>& b, int block = 256) {
  if (n <= 0 || nb <= 0) return;
  if (nb == 1) { parallel_for(s, n, b.f[0], block); return; }
  const int64_t grid = (n + block - 1) / block;
  ++launch_counter();
  pf_batch_kernel<F><<<dim3((unsigned)grid, (unsigned)nb), block, 0, s>>>(n, b);
  KNP_CUDA(cudaGetLastError());
}
#endif

}  // namespace knp
