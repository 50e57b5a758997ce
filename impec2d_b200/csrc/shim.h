// shim.h -- one kernel source, two builds.
//
//   nvcc (default)      : real CUDA for sm_100a.  This is the product.
//   g++ -DIMPEC_EMU     : the same kernels executed by a small CPU emulator (emu.cpp) so that index
//                         logic can be checked against the oracle in this GPU-less container BEFORE
//                         spending B200 time.  The emulated library is test infrastructure: it is
//                         built as libimpec2d_emu.so, only tests/ load it, and the product loader
//                         (impec2d_b200/_lib.py) refuses to fall back to it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>

#ifndef IMPEC_EMU
// ------------------------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#define IMPEC_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define IMPEC_LAUNCH_COOP(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define IMPEC_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; type *name = reinterpret_cast<type *>(name##_raw_)
#define IMPEC_SHFL_DOWN(v, d) __shfl_down_sync(0xffffffffu, (v), (d))
#define IMPEC_SHFL_UP(v, d) __shfl_up_sync(0xffffffffu, (v), (d))
#define IMPEC_SHFL_IDX(v, l) __shfl_sync(0xffffffffu, (v), (l))
#define IMPEC_SYNCTHREADS() __syncthreads()
#define IMPEC_SYNCWARP() __syncwarp()
#define IMPEC_HD __host__ __device__ __forceinline__
#define IMPEC_D __device__ __forceinline__
// 16-byte asynchronous global->shared copy (LDGSTS, bypasses L1) and its completion wait
__device__ __forceinline__ void impec_cp_async16(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void impec_cp_async_wait_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;\n" ::: "memory"); }
#else
// ------------------------------------------------------------------------------------ emulator
#include <functional>
#include <cstdlib>
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
namespace emu {
extern dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
extern unsigned char *dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, bool coop, const std::function<void()> &body);
void syncthreads();
double shfl(double v, int src_lane_delta, int mode);   // mode 0 down, 1 up, 2 idx
}  // namespace emu
#define threadIdx emu::threadIdx_
#define blockIdx emu::blockIdx_
#define blockDim emu::blockDim_
#define gridDim emu::gridDim_
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define IMPEC_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch((grid), (block), (smem), false, [&]() { kernel(__VA_ARGS__); })
#define IMPEC_LAUNCH_COOP(kernel, grid, block, smem, stream, ...) emu::launch((grid), (block), (smem), true, [&]() { kernel(__VA_ARGS__); })
#define IMPEC_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu::dyn_smem)
#define IMPEC_SHFL_DOWN(v, d) emu::shfl((v), (d), 0)
#define IMPEC_SHFL_UP(v, d) emu::shfl((v), (d), 1)
#define IMPEC_SHFL_IDX(v, l) emu::shfl((v), (l), 2)
#define IMPEC_SYNCTHREADS() emu::syncthreads()
#define IMPEC_SYNCWARP() emu::syncthreads()   /* block barrier: callers keep all 32 lanes of the warp on the same path */
#define IMPEC_HD inline
#define IMPEC_D inline
inline void impec_cp_async16(void *smem_dst, const void *gsrc) { std::memcpy(smem_dst, gsrc, 16); }
inline void impec_cp_async_wait_all() {}
inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) std::memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return 0;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
enum { cudaStreamNonBlocking = 1 };
#endif
