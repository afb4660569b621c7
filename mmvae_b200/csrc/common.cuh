// Shared helpers for the cmmvae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/cmmvae_b200.h"

namespace cmmvae {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

#define CMMVAE_REQUIRE(cond, ...)  \
  do {                             \
    if (!(cond)) {                 \
      cmmvae::set_error(__VA_ARGS__); \
      return -1;                   \
    }                              \
  } while (0)

// Programmatic dependent launch: a kernel launched through launch_pdl may be scheduled while its predecessor in
// the stream is still running; it must execute pdl_sync() before it touches global memory (the call returns
// once the predecessor grid has completed and its writes are visible).  Everything before pdl_sync() -- barrier
// init, TMEM allocation, tensor-map prefetch, block scheduling itself -- overlaps the predecessor's tail.
// pdl_sync() also lets this grid's own successor start its prologue.  CMMVAE_PDL=0 turns the attribute off.
bool pdl_enabled();
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// short kernels: let the successor in right away.  The long persistent tensor-pipe kernels instead call
// pdl_wait() after their prologue and pdl_trigger() once their last unit's MMAs are issued, so a successor's
// CTAs do not sit on the SMs for the whole run
__device__ __forceinline__ void pdl_sync() {
  pdl_trigger();
  pdl_wait();
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

constexpr int kNumSMs = 148;
// SMs the persistent kernels may plan for (grid sizes, split-K factors).  148 by default; the host lowers it
// while NCCL kernels share the GPU so that every planned CTA is resident at once (cmmvae_set_sm_budget).
int sm_budget();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in thread 0
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem32) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem32[threadIdx.x] : T(0);
  if (w == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// counter-based Bernoulli keep mask for dropout: same bits in forward and backward
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (uint32_t)x;
}
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx, float p) {
  uint32_t r = mix32(seed * 0x9E3779B97F4A7C15ULL + idx);
  return (r >> 8) * (1.0f / 16777216.0f) >= p;
}

}  // namespace cmmvae
