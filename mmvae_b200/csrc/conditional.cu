// Conditional layers on the latent: ConditionalLayer.forward (components.py:369-413) routes every
// cell through the FCBlock of ITS metadata value -- one Linear(Zin, Zout) [+ LayerNorm without affine, components.py:
// 277-278] [+ ReLU] per value (the shipped topology: configs/model/human_only.yaml:61-68) -- and ConditionalLayers
// (components.py:581-631) chains the batch keys or concatenates their outputs.  The reference does this with a Python
// dict of row lists and one index_select / module call / index_copy_ per value; here the host sorts the rows of a
// batch by value once (tiles of <= 32 rows that share a value = a "slot" of the parameter bank) and ONE launch per
// direction handles every value of every batch key:
//   forward   y = LN(x W_s^T + b_s) for the rows of each tile            (CUDA cores, fp32: 0.3 GFLOP per step)
//   backward  LayerNorm backward, dx = dy W_s, dW_s += dy^T x, db_s += sum dy (atomics into the slot's gradient)
// Only the slots present in the batch are touched afterwards: their gradients are zeroed before, added to the clip
// norm and stepped by Adam with THEIR OWN step count -- torch.optim.Adam skips parameters whose grad is None
// (unused modules after zero_grad(set_to_none=True)): no decay, no moment update, no step for them.
#include "common.cuh"

namespace cmmvae {

constexpr int kCondRows = 32;       // rows per tile
constexpr int kCondThreads = 128;

// dynamic smem: xsT[Zin][32] | ysT[Zout][32]   (row index fastest: one LDS.128 = 4 rows of one column)
__global__ void __launch_bounds__(kCondThreads)
cond_fwd_kernel(const float* __restrict__ params, long long S, int Zin, int Zout, const int4* __restrict__ tiles,
                const int* __restrict__ rows, const float* __restrict__ x, int ldx, float* __restrict__ out,
                __nv_bfloat16* __restrict__ out16, float* __restrict__ pre, int ldo, const int* __restrict__ ooff,
                float* __restrict__ rstd, int B, int layer_norm, int relu) {
  extern __shared__ float smem[];
  float* xsT = smem;
  float* ysT = smem + (size_t)Zin * kCondRows;
  __shared__ int row_s[kCondRows];
  pdl_sync();
  const int4 t = tiles[blockIdx.x];
  const int slot = t.x, start = t.y, count = t.z, cond = t.w;
  const float* W = params + (long long)slot * S;
  const float* bias = W + (long long)Zout * Zin;
  const int tid = threadIdx.x;
  if (tid < kCondRows) row_s[tid] = tid < count ? rows[start + tid] : -1;
  __syncthreads();
  for (int i = tid; i < Zin * kCondRows; i += kCondThreads) {
    const int r = i / Zin, k = i - r * Zin;            // coalesced along k
    const int row = row_s[r];
    xsT[k * kCondRows + r] = row >= 0 ? x[(long long)row * ldx + k] : 0.f;
  }
  __syncthreads();
  for (int j = tid; j < Zout; j += kCondThreads) {
    float acc[kCondRows];
    const float bj = bias[j];
#pragma unroll
    for (int r = 0; r < kCondRows; ++r) acc[r] = bj;
    const float* Wj = W + (long long)j * Zin;
    for (int k = 0; k < Zin; ++k) {
      const float w = __ldg(Wj + k);
      const float4* xr = reinterpret_cast<const float4*>(xsT + k * kCondRows);
#pragma unroll
      for (int q = 0; q < kCondRows / 4; ++q) {
        const float4 v = xr[q];
        acc[4 * q] = fmaf(w, v.x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(w, v.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(w, v.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(w, v.w, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int r = 0; r < kCondRows; ++r) ysT[j * kCondRows + r] = acc[r];
  }
  __syncthreads();
  // LayerNorm (no affine, eps 1e-5, biased variance) + activation: one warp per row
  const int warp = tid >> 5, lane = tid & 31;
  const int col0 = ooff[cond];
  for (int r = warp; r < count; r += kCondThreads / 32) {
    const int row = row_s[r];
    float mean = 0.f, rs = 1.f;
    if (layer_norm) {
      float s = 0.f;
      for (int j = lane; j < Zout; j += 32) s += ysT[j * kCondRows + r];
      s = warp_sum(s);
      mean = s / (float)Zout;
      float q = 0.f;
      for (int j = lane; j < Zout; j += 32) {
        const float d = ysT[j * kCondRows + r] - mean;
        q = fmaf(d, d, q);
      }
      q = warp_sum(q);
      rs = rsqrtf(q / (float)Zout + 1e-5f);
      if (lane == 0) rstd[(long long)cond * B + row] = rs;
    }
    for (int j = lane; j < Zout; j += 32) {
      const float h = (ysT[j * kCondRows + r] - mean) * rs;
      const float o = relu ? fmaxf(h, 0.f) : h;
      const long long at = (long long)row * ldo + col0 + j;
      pre[at] = h;
      out[at] = o;
      if (out16) out16[at] = __float2bfloat16(o);
    }
  }
}

// dynamic smem: xs[32][Zin] (row major: thread k reads a column into registers) | dysT[Zout][32]
__global__ void __launch_bounds__(kCondThreads)
cond_bwd_kernel(const float* __restrict__ params, float* __restrict__ grads, long long S, int Zin, int Zout,
                const int4* __restrict__ tiles, const int* __restrict__ rows, const float* __restrict__ x, int ldx,
                const float* __restrict__ dout, const float* __restrict__ pre, int ldo,
                const int* __restrict__ ooff, const float* __restrict__ rstd, int B, float* __restrict__ dx,
                int lddx, const int* __restrict__ dxoff, int layer_norm, int relu) {
  extern __shared__ float smem[];
  float* xs = smem;
  float* dysT = smem + (size_t)Zin * kCondRows;
  __shared__ int row_s[kCondRows];
  pdl_sync();
  const int4 t = tiles[blockIdx.x];
  const int slot = t.x, start = t.y, count = t.z, cond = t.w;
  const float* W = params + (long long)slot * S;
  float* gW = grads + (long long)slot * S;
  float* gb = gW + (long long)Zout * Zin;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < kCondRows) row_s[tid] = tid < count ? rows[start + tid] : -1;
  __syncthreads();
  for (int i = tid; i < Zin * kCondRows; i += kCondThreads) {
    const int r = i / Zin, k = i - r * Zin;
    const int row = row_s[r];
    xs[r * Zin + k] = row >= 0 ? x[(long long)row * ldx + k] : 0.f;
  }
  // dy of the tile's rows (LayerNorm backward without affine: dy = rstd (g - mean(g) - h mean(g h)))
  const int col0 = ooff[cond];
  for (int r = warp; r < kCondRows; r += kCondThreads / 32) {
    const int row = row_s[r];
    if (row < 0) {
      for (int j = lane; j < Zout; j += 32) dysT[j * kCondRows + r] = 0.f;
      continue;
    }
    const long long at = (long long)row * ldo + col0;
    float m1 = 0.f, m2 = 0.f, rs = 1.f;
    if (layer_norm) {
      for (int j = lane; j < Zout; j += 32) {
        const float h = pre[at + j];
        const float g = (relu && h <= 0.f) ? 0.f : dout[at + j];
        m1 += g;
        m2 = fmaf(g, h, m2);
      }
      m1 = warp_sum(m1) / (float)Zout;
      m2 = warp_sum(m2) / (float)Zout;
      rs = rstd[(long long)cond * B + row];
    }
    for (int j = lane; j < Zout; j += 32) {
      const float h = pre[at + j];
      const float g = (relu && h <= 0.f) ? 0.f : dout[at + j];
      dysT[j * kCondRows + r] = layer_norm ? rs * (g - m1 - h * m2) : g;
    }
  }
  __syncthreads();
  // db_s[j] += sum_r dy[r][j]
  for (int j = tid; j < Zout; j += kCondThreads) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kCondRows; ++r) s += dysT[j * kCondRows + r];
    atomicAdd(gb + j, s);
  }
  const int xoff = dxoff[cond];
  for (int k = tid; k < Zin; k += kCondThreads) {
    // dx[r][k] = sum_j dy[r][j] W[j][k]   (W read coalesced along k)
    float acc[kCondRows];
#pragma unroll
    for (int r = 0; r < kCondRows; ++r) acc[r] = 0.f;
    for (int j = 0; j < Zout; ++j) {
      const float w = __ldg(W + (long long)j * Zin + k);
      const float4* dr = reinterpret_cast<const float4*>(dysT + j * kCondRows);
#pragma unroll
      for (int q = 0; q < kCondRows / 4; ++q) {
        const float4 v = dr[q];
        acc[4 * q] = fmaf(w, v.x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(w, v.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(w, v.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(w, v.w, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int r = 0; r < kCondRows; ++r) {
      const int row = row_s[r];
      if (row >= 0) dx[(long long)row * lddx + xoff + k] = acc[r];
    }
    // dW_s[j][k] += sum_r dy[r][j] x[r][k]   (column k of the tile's inputs in registers)
    float xr[kCondRows];
#pragma unroll
    for (int r = 0; r < kCondRows; ++r) xr[r] = xs[r * Zin + k];
    for (int j = 0; j < Zout; ++j) {
      const float4* dr = reinterpret_cast<const float4*>(dysT + j * kCondRows);
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < kCondRows / 4; ++q) {
        const float4 v = dr[q];
        s = fmaf(v.x, xr[4 * q], s);
        s = fmaf(v.y, xr[4 * q + 1], s);
        s = fmaf(v.z, xr[4 * q + 2], s);
        s = fmaf(v.w, xr[4 * q + 3], s);
      }
      atomicAdd(gW + (long long)j * Zin + k, s);
    }
  }
}

__global__ void cond_zero_kernel(float* __restrict__ grads, long long S, const int* __restrict__ present) {
  pdl_sync();
  float4* g = reinterpret_cast<float4*>(grads + (long long)present[blockIdx.x] * S);
  for (long long i = blockIdx.y * (long long)blockDim.x + threadIdx.x; i < S / 4; i += (long long)gridDim.y * blockDim.x)
    g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void cond_sumsq_kernel(const float* __restrict__ grads, long long S, const int* __restrict__ present,
                                  double* __restrict__ out) {
  pdl_sync();
  const float4* g = reinterpret_cast<const float4*>(grads + (long long)present[blockIdx.x] * S);
  float s = 0.f;
  for (long long i = blockIdx.y * (long long)blockDim.x + threadIdx.x; i < S / 4; i += (long long)gridDim.y * blockDim.x) {
    const float4 v = g[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  __shared__ float part[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, (double)v);
  }
}

// torch.optim.Adam on the present slots, each with its own step count t = steps[slot] + 1 (the same clip
// coefficient as the rest of the optimizer group: the norm covers the group's dense part AND these slots)
__global__ void cond_adam_kernel(float* __restrict__ params, const float* __restrict__ grads, float* __restrict__ m,
                                 float* __restrict__ v, long long S, const int* __restrict__ present,
                                 const int* __restrict__ steps, const double* __restrict__ norm_sq, float max_norm,
                                 float grad_scale, float lr, double b1, double b2, float eps, float wd) {
  pdl_sync();
  const int slot = present[blockIdx.x];
  const double t = (double)(steps[slot] + 1);
  const float bc1 = (float)(1.0 - pow(b1, t));
  const float isb2 = (float)(1.0 / sqrt(1.0 - pow(b2, t)));
  float coef = 1.f;
  if (max_norm > 0.f && norm_sq) {
    const float total = (float)sqrt(*norm_sq) * fabsf(grad_scale);
    coef = fminf(1.f, max_norm / (total + 1e-6f));
  }
  const float gs = coef * grad_scale;
  const float fb1 = (float)b1, fb2 = (float)b2;
  const long long base = (long long)slot * S;
  for (long long i = blockIdx.y * (long long)blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.y * blockDim.x) {
    float p = params[base + i], mm = m[base + i], vv = v[base + i];
    float g = grads[base + i] * gs;
    g = fmaf(wd, p, g);
    mm = mm + (g - mm) * (1.f - fb1);
    vv = vv * fb2 + (1.f - fb2) * g * g;
    const float denom = sqrtf(vv) * isb2 + eps;
    p = p - (lr / bc1) * (mm / denom);
    params[base + i] = p; m[base + i] = mm; v[base + i] = vv;
  }
}

__global__ void cond_step_inc_kernel(int* __restrict__ steps, const int* __restrict__ present, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) steps[present[i]] += 1;
}
}  // namespace cmmvae

using namespace cmmvae;

static size_t cond_smem(int Zin, int Zout) { return (size_t)(Zin + Zout) * kCondRows * sizeof(float); }

extern "C" int cmmvae_cond_fwd(const float* params, long long slot_stride, int Zin, int Zout, const int32_t* tiles,
                               int n_tiles, const int32_t* rows, const float* x, int ldx, float* out, void* out_bf16,
                               float* pre, int ldo, const int32_t* out_col, float* rstd, int B, int layer_norm, int relu,
                               void* stream) {
  if (n_tiles <= 0) return 0;
  CMMVAE_REQUIRE(params && tiles && rows && x && out && pre && out_col && rstd, "cond_fwd: null pointer");
  CMMVAE_REQUIRE(Zin > 0 && Zout > 0 && slot_stride % 4 == 0 && slot_stride >= (long long)Zout * Zin + Zout,
                 "cond_fwd: bad sizes");
  const size_t smem = cond_smem(Zin, Zout);
  CMMVAE_REQUIRE(smem <= 200 * 1024, "cond_fwd: Zin + Zout must be <= 1600");
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaFuncSetAttribute(cond_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  launch_pdl(cond_fwd_kernel, dim3(n_tiles), dim3(kCondThreads), smem, (cudaStream_t)stream, params, slot_stride, Zin,
             Zout, (const int4*)tiles, rows, x, ldx, out, (__nv_bfloat16*)out_bf16, pre, ldo, out_col, rstd, B,
             layer_norm, relu);
  return check_launch("cond_fwd");
}

extern "C" int cmmvae_cond_bwd(const float* params, float* grads, long long slot_stride, int Zin, int Zout,
                               const int32_t* tiles, int n_tiles, const int32_t* rows, const float* x, int ldx,
                               const float* dout, const float* pre, int ldo, const int32_t* out_col, const float* rstd,
                               int B, float* dx, int lddx, const int32_t* dx_col, int layer_norm, int relu,
                               void* stream) {
  if (n_tiles <= 0) return 0;
  CMMVAE_REQUIRE(params && grads && tiles && rows && x && dout && pre && out_col && rstd && dx && dx_col,
                 "cond_bwd: null pointer");
  CMMVAE_REQUIRE(Zin > 0 && Zout > 0 && slot_stride % 4 == 0, "cond_bwd: bad sizes");
  const size_t smem = cond_smem(Zin, Zout);
  CMMVAE_REQUIRE(smem <= 200 * 1024, "cond_bwd: Zin + Zout must be <= 1600");
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaFuncSetAttribute(cond_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  launch_pdl(cond_bwd_kernel, dim3(n_tiles), dim3(kCondThreads), smem, (cudaStream_t)stream, params, grads,
             slot_stride, Zin, Zout, (const int4*)tiles, rows, x, ldx, dout, pre, ldo, out_col, rstd, B, dx, lddx,
             dx_col, layer_norm, relu);
  return check_launch("cond_bwd");
}

extern "C" int cmmvae_cond_zero_grads(float* grads, long long slot_stride, const int32_t* present, int n_present,
                                      void* stream) {
  if (n_present <= 0) return 0;
  CMMVAE_REQUIRE(grads && present && slot_stride % 4 == 0, "cond_zero_grads: bad arguments");
  const int chunks = (int)((slot_stride / 4 + 255) / 256);
  launch_pdl(cond_zero_kernel, dim3(n_present, chunks < 64 ? chunks : 64), dim3(256), 0, (cudaStream_t)stream, grads,
             slot_stride, present);
  return check_launch("cond_zero_grads");
}

extern "C" int cmmvae_cond_sumsq(const float* grads, long long slot_stride, const int32_t* present, int n_present,
                                 double* out, void* stream) {
  if (n_present <= 0) return 0;
  CMMVAE_REQUIRE(grads && present && out && slot_stride % 4 == 0, "cond_sumsq: bad arguments");
  const int chunks = (int)((slot_stride / 4 + 255) / 256);
  launch_pdl(cond_sumsq_kernel, dim3(n_present, chunks < 16 ? chunks : 16), dim3(256), 0, (cudaStream_t)stream, grads,
             slot_stride, present, out);
  return check_launch("cond_sumsq");
}

extern "C" int cmmvae_cond_adam(float* params, const float* grads, float* m, float* v, long long slot_stride,
                                const int32_t* present, int n_present, int32_t* steps, const double* norm_sq,
                                float max_norm, float grad_scale, float lr, double beta1, double beta2, float eps,
                                float weight_decay, void* stream) {
  if (n_present <= 0) return 0;
  CMMVAE_REQUIRE(params && grads && m && v && present && steps, "cond_adam: null pointer");
  const int chunks = (int)((slot_stride + 1023) / 1024);
  // (the betas travel as doubles: the bias corrections 1 - beta ** t are taken as Python takes them)
  launch_pdl(cond_adam_kernel, dim3(n_present, chunks < 64 ? chunks : 64), dim3(256), 0, (cudaStream_t)stream, params,
             grads, m, v, slot_stride, present, (const int*)steps, norm_sq, max_norm, grad_scale, lr, beta1, beta2,
             eps, weight_decay);
  int rc = check_launch("cond_adam");
  if (rc) return rc;
  launch_pdl(cond_step_inc_kernel, dim3((n_present + 127) / 128), dim3(128), 0, (cudaStream_t)stream, steps, present,
             n_present);
  return check_launch("cond_step_inc");
}
