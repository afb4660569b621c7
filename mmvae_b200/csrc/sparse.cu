// Sparse (CSR) kernels of the CMMVAE step: the expert-encoder first layer (K1), its weight
// gradient (K1b), the CSR->CSC transposition feeding it, and the unfused ReLU + sum-MSE against
// the CSR batch (K6/K7 without the to_dense round trip).
#include "common.cuh"

namespace cmmvae {

// ---------------------------------------------------------------------------------------------
// K1: Y[b, :] = sum_j val[j] * Wt[col[j], :] + bias      (one CTA per (row, column block))
// Each thread owns V consecutive output columns and streams the row's non-zeros through a
// shared-memory staging buffer of (col, val) pairs; weight rows are read as one coalesced,
// vectorised segment per non-zero (a warp reads 32*V contiguous elements).
// ---------------------------------------------------------------------------------------------
constexpr int kStage = 256;

template <typename WT, int V>
struct WVec;
template <>
struct WVec<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&w)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
  }
};
template <>
struct WVec<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&w)[1]) { w[0] = __ldg(p); }
};
template <>
struct WVec<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&w)[8]) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    w[0] = bf16_lo(t.x); w[1] = bf16_hi(t.x); w[2] = bf16_lo(t.y); w[3] = bf16_hi(t.y);
    w[4] = bf16_lo(t.z); w[5] = bf16_hi(t.z); w[6] = bf16_lo(t.w); w[7] = bf16_hi(t.w);
  }
};
template <>
struct WVec<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&w)[1]) {
    w[0] = __bfloat162float(*p);
  }
};

template <typename WT, int V, int UNROLL>
__global__ void __launch_bounds__(256) csr_linear_fwd_kernel(const int32_t* __restrict__ crow,
                                                             const int32_t* __restrict__ col,
                                                             const float* __restrict__ val, int H,
                                                             const WT* __restrict__ Wt,
                                                             const float* __restrict__ bias,
                                                             float* __restrict__ Y) {
  __shared__ int32_t s_col[kStage];
  __shared__ float s_val[kStage];
  const int row = blockIdx.x;
  const int h0 = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  const bool active = h0 < H;
  const int start = crow[row], end = crow[row + 1];
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  const WT* wbase = Wt + h0;

  for (int base = start; base < end; base += kStage) {
    const int n = min(kStage, end - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      s_col[i] = col[base + i];
      s_val[i] = val[base + i];
    }
    __syncthreads();
    if (active) {
      int j = 0;
      for (; j + UNROLL <= n; j += UNROLL) {
        float w[UNROLL][V];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) WVec<WT, V>::load(wbase + (size_t)s_col[j + u] * H, w[u]);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const float x = s_val[j + u];
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = fmaf(x, w[u][v], acc[v]);
        }
      }
      for (; j < n; ++j) {
        float w[V];
        WVec<WT, V>::load(wbase + (size_t)s_col[j] * H, w);
        const float x = s_val[j];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = fmaf(x, w[v], acc[v]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int v = 0; v < V; ++v) Y[(size_t)row * H + h0 + v] = acc[v] + (bias ? bias[h0 + v] : 0.f);
  }
}

template <typename WT, int V>
static int launch_csr_fwd(const int32_t* crow, const int32_t* col, const float* val, int B, int H,
                          const void* Wt, const float* bias, float* Y, cudaStream_t st) {
  int threads_needed = (H + V - 1) / V;
  int block = min(256, ((threads_needed + 31) / 32) * 32);
  dim3 grid(B, (threads_needed + block - 1) / block);
  csr_linear_fwd_kernel<WT, V, 8><<<grid, block, 0, st>>>(crow, col, val, H, (const WT*)Wt, bias, Y);
  return check_launch("csr_linear_fwd");
}

// ---------------------------------------------------------------------------------------------
// CSR -> CSC
// ---------------------------------------------------------------------------------------------
__global__ void csc_hist_kernel(const int32_t* __restrict__ col, long long nnz, int32_t* __restrict__ cptr) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
       i += (long long)gridDim.x * blockDim.x)
    atomicAdd(&cptr[col[i] + 1], 1);
}

// single-CTA in-place inclusive scan of a[0..n) (a[0] == 0 on entry => exclusive pointers);
// coalesced: 1024 consecutive elements per iteration, warp-shuffle scan + carried offset
__global__ void __launch_bounds__(1024) scan_kernel(int32_t* __restrict__ a, int n, int32_t* __restrict__ copy) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + t;
    int32_t v = (i < n) ? a[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    if (w == 0) {
      int32_t x = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += u;
      }
      warp_tot[lane] = x;
    }
    __syncthreads();
    const int32_t carry = carry_s;
    v += carry + (w > 0 ? warp_tot[w - 1] : 0);
    if (i < n) {
      a[i] = v;
      copy[i] = v;
    }
    __syncthreads();
    if (t == 1023) carry_s = v;
    __syncthreads();
  }
}

__global__ void csc_scatter_kernel(const int32_t* __restrict__ crow, const int32_t* __restrict__ col,
                                   const float* __restrict__ val, int B, int32_t* __restrict__ cursor,
                                   int32_t* __restrict__ ridx, float* __restrict__ cval) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < B; r += nwarps) {
    const int s = crow[r], e = crow[r + 1];
    for (int j = s + lane; j < e; j += 32) {
      const int pos = atomicAdd(&cursor[col[j]], 1);
      ridx[pos] = r;
      cval[pos] = val[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1b: dWt[g, :] = sum_{i in column g} cval[i] * dY[ridx[i], :]   (gather; every gene row is
// written exactly once, genes without entries get exact zeros)
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) csr_linear_bwd_w_kernel(const int32_t* __restrict__ cptr,
                                                               const int32_t* __restrict__ ridx,
                                                               const float* __restrict__ cval, int G, int H,
                                                               const float* __restrict__ dY,
                                                               float* __restrict__ dWt) {
  const int h0 = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (h0 >= H) return;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int s = cptr[g], e = cptr[g + 1];
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    int j = s;
    for (; j + 4 <= e; j += 4) {
      float w[4][V];
      float x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        x[u] = __ldg(cval + j + u);
        WVec<float, V>::load(dY + (size_t)__ldg(ridx + j + u) * H + h0, w[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = fmaf(x[u], w[u][v], acc[v]);
    }
    for (; j < e; ++j) {
      float w[V];
      const float x = __ldg(cval + j);
      WVec<float, V>::load(dY + (size_t)__ldg(ridx + j) * H + h0, w);
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = fmaf(x, w[v], acc[v]);
    }
    if (V == 4) {
      *reinterpret_cast<float4*>(dWt + (size_t)g * H + h0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) dWt[(size_t)g * H + h0 + v] = acc[v];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// unfused ReLU + sum-MSE against the CSR batch (one CTA per cell)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mse_relu_csr_kernel(float* __restrict__ logits, int ldl, int G,
                                                           const int32_t* __restrict__ crow,
                                                           const int32_t* __restrict__ col,
                                                           const float* __restrict__ val, int write_xhat,
                                                           float* __restrict__ dl32, __nv_bfloat16* __restrict__ dl16,
                                                           int ldd, double* __restrict__ loss_sum) {
  __shared__ double red[32];
  const int b = blockIdx.x;
  float* lrow = logits + (size_t)b * ldl;
  float part = 0.f;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float y = lrow[g];
    const float xh = fmaxf(y, 0.f);
    part = fmaf(xh, xh, part);
    if (write_xhat) lrow[g] = xh;
    if (dl32) dl32[(size_t)b * ldd + g] = 2.f * xh;
    if (dl16) dl16[(size_t)b * ldd + g] = __float2bfloat16(2.f * xh);
  }
  __syncthreads();  // the sparse fix-up below re-reads / overwrites what the dense pass wrote
  const int s = crow[b], e = crow[b + 1];
  for (int j = s + threadIdx.x; j < e; j += blockDim.x) {
    const int c = col[j];
    const float x = val[j];
    const float xh = fmaxf(lrow[c], 0.f);
    part += x * x - 2.f * x * xh;
    const float d = xh > 0.f ? 2.f * (xh - x) : 0.f;
    if (dl32) dl32[(size_t)b * ldd + c] = d;
    if (dl16) dl16[(size_t)b * ldd + c] = __float2bfloat16(d);
  }
  double tot = block_sum<double>((double)part, red);
  if (threadIdx.x == 0) atomicAdd(loss_sum, tot);
}

}  // namespace cmmvae

using namespace cmmvae;

extern "C" int cmmvae_csr_linear_fwd(const int32_t* crow, const int32_t* col, const float* val, int B, int G,
                                     int H, const void* Wt, int w_dtype, const float* bias, float* Y,
                                     void* stream) {
  CMMVAE_REQUIRE(B > 0 && G > 0 && H > 0, "csr_linear_fwd: bad shape B=%d G=%d H=%d", B, G, H);
  cudaStream_t st = (cudaStream_t)stream;
  if (w_dtype == CMMVAE_F32) {
    if (H % 4 == 0) return launch_csr_fwd<float, 4>(crow, col, val, B, H, Wt, bias, Y, st);
    return launch_csr_fwd<float, 1>(crow, col, val, B, H, Wt, bias, Y, st);
  } else if (w_dtype == CMMVAE_BF16) {
    if (H % 8 == 0) return launch_csr_fwd<__nv_bfloat16, 8>(crow, col, val, B, H, Wt, bias, Y, st);
    return launch_csr_fwd<__nv_bfloat16, 1>(crow, col, val, B, H, Wt, bias, Y, st);
  }
  set_error("csr_linear_fwd: unknown w_dtype %d", w_dtype);
  return -1;
}

extern "C" int cmmvae_csr_transpose(const int32_t* crow, const int32_t* col, const float* val, int B, int G,
                                    long long nnz, int32_t* cptr, int32_t* ridx, float* cval, int32_t* cursor,
                                    void* stream) {
  CMMVAE_REQUIRE(B > 0 && G > 0 && nnz >= 0, "csr_transpose: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(cptr, 0, sizeof(int32_t) * (size_t)(G + 1), st);
  if (nnz > 0) {
    long long want = (nnz + 255) / 256; int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    csc_hist_kernel<<<blocks, 256, 0, st>>>(col, nnz, cptr);
    if (int rc = check_launch("csc_hist")) return rc;
  }
  scan_kernel<<<1, 1024, 0, st>>>(cptr, G + 1, cursor);
  if (int rc = check_launch("csc_scan")) return rc;
  if (nnz > 0) {
    int blocks = min((B + 7) / 8, 148 * 8);
    csc_scatter_kernel<<<blocks, 256, 0, st>>>(crow, col, val, B, cursor, ridx, cval);
    if (int rc = check_launch("csc_scatter")) return rc;
  }
  return 0;
}

extern "C" int cmmvae_csr_linear_bwd_w(const int32_t* cptr, const int32_t* ridx, const float* cval, int B, int G,
                                       int H, const float* dY, float* dWt, void* stream) {
  CMMVAE_REQUIRE(B > 0 && G > 0 && H > 0, "csr_linear_bwd_w: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int V = (H % 4 == 0) ? 4 : 1;
  int threads_needed = (H + V - 1) / V;
  int block = min(256, ((threads_needed + 31) / 32) * 32);
  dim3 grid(min(G, 148 * 32), (threads_needed + block - 1) / block);
  if (V == 4)
    csr_linear_bwd_w_kernel<4><<<grid, block, 0, st>>>(cptr, ridx, cval, G, H, dY, dWt);
  else
    csr_linear_bwd_w_kernel<1><<<grid, block, 0, st>>>(cptr, ridx, cval, G, H, dY, dWt);
  return check_launch("csr_linear_bwd_w");
}

extern "C" int cmmvae_mse_relu_csr(float* logits, int ldl, int B, int G, const int32_t* crow, const int32_t* col,
                                   const float* val, int write_xhat, float* dlogits_f32, void* dlogits_bf16,
                                   int ldd, double* loss_sum, void* stream) {
  CMMVAE_REQUIRE(B > 0 && G > 0 && ldl >= G, "mse_relu_csr: bad shape");
  CMMVAE_REQUIRE((!dlogits_f32 && !dlogits_bf16) || ldd >= G, "mse_relu_csr: ldd < G");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(loss_sum, 0, sizeof(double), st);
  mse_relu_csr_kernel<<<B, 256, 0, st>>>(logits, ldl, G, crow, col, val, write_xhat, dlogits_f32,
                                          (__nv_bfloat16*)dlogits_bf16, ldd, loss_sum);
  return check_launch("mse_relu_csr");
}
