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

// one torch.optim.Adam update (same arithmetic as clip_adam_kernel in dense_basic.cu)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float gscale, float lr, float b1,
                                         float b2, float eps, float wd, float bc1, float inv_sqrt_bc2) {
  g = g * gscale;
  g = fmaf(wd, p, g);
  m = m + (g - m) * (1.f - b1);
  v = v * b2 + (1.f - b2) * g * g;
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
  p = p - (lr / bc1) * (m / denom);
}

// Tile kernels.  A block of 128 threads = 4 warps owns one tile (<= 32 rows of one slot).  Register tiling: thread
// (tx = lane, ty = warp) computes rows 8 ty .. 8 ty + 7 x columns {tx, tx + 32, tx + 64, tx + 96} of a 32 x 128 output
// block; operands sit in shared memory row-major with a pitch of P = K + 4 floats (K <= 128 per pass), which makes
// the lanes' 16-byte reads of four consecutive k conflict free (pitch = 4 mod 32 words) and every global row a
// 512-byte coalesced load.  Warps whose 8 rows lie beyond the tile's row count skip the arithmetic: most tiles of a
// many-valued key (donor_id) hold one or two cells.
constexpr int kCondK = 128;              // K handled per pass
constexpr int kCondP = kCondK + 4;       // shared-memory pitch

// one row piece of 4 floats at column k (global), zero beyond K; 16-byte load when the row pitch allows it
__device__ __forceinline__ float4 cond_load4(const float* __restrict__ s, int k, int K, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec && k + 3 < K) return __ldg(reinterpret_cast<const float4*>(s));
  if (k < K) v.x = __ldg(s);
  if (k + 1 < K) v.y = __ldg(s + 1);
  if (k + 2 < K) v.z = __ldg(s + 2);
  if (k + 3 < K) v.w = __ldg(s + 3);
  return v;
}

// rows x K block of a row-major global matrix (ld) -> smem [rows][kCondP].  Eight independent 512-byte row loads
// are issued per warp before the first one is stored (a load-store pair per iteration would serialise on the
// memory latency: 32 round trips per block).
__device__ __forceinline__ void cond_stage_rows(float* __restrict__ dst, const float* __restrict__ src, long long ld,
                                                int n_rows, int k0, int K, int warp, int lane) {
  const bool vec = (ld & 3) == 0;
  const int k = 4 * lane;
  for (int r0 = warp; r0 < n_rows; r0 += 8 * (kCondThreads / 32)) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + u * (kCondThreads / 32);
      v[u] = r < n_rows ? cond_load4(src + (long long)r * ld + k0 + k, k0 + k, K, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + u * (kCondThreads / 32);
      if (r < n_rows) *reinterpret_cast<float4*>(dst + r * kCondP + k) = v[u];
    }
  }
}

// gathered rows of x (row list in row_s, -1 = none) -> smem [32][kCondP]: a warp's 8 rows are loaded together
__device__ __forceinline__ void cond_stage_x(float* __restrict__ dst, const float* __restrict__ x, long long ldx,
                                             const int* row_s, int k0, int K, int warp, int lane,
                                             int pitch = kCondP, bool pad_chunk = true) {
  const bool vec = ((ldx | k0) & 3) == 0;
  const int k = 4 * lane;
  float4 v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int row = row_s[warp + u * (kCondThreads / 32)];
    v[u] = row >= 0 ? cond_load4(x + (long long)row * ldx + k0 + k, k0 + k, K, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (!pad_chunk && k0 + k >= K) return;      // (pitch K + 4: nothing exists beyond column K + 3)
#pragma unroll
  for (int u = 0; u < 8; ++u)
    *reinterpret_cast<float4*>(dst + (warp + u * (kCondThreads / 32)) * pitch + k0 + k) = v[u];
}

// acc[i][c] += sum_k A[8 ty + i][k] * Bm[tx + 32 c][k]   (A, Bm: smem [.][kCondP], K = kCondK)
__device__ __forceinline__ void cond_mma_nt(float (&acc)[8][4], const float* __restrict__ A,
                                            const float* __restrict__ Bm, int ty, int tx) {
#pragma unroll 2
  for (int k = 0; k < kCondK; k += 4) {
    float4 b[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(Bm + (tx + 32 * c) * kCondP + k);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(A + (8 * ty + i) * kCondP + k);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[i][c] = fmaf(a.x, b[c].x, acc[i][c]);
        acc[i][c] = fmaf(a.y, b[c].y, acc[i][c]);
        acc[i][c] = fmaf(a.z, b[c].z, acc[i][c]);
        acc[i][c] = fmaf(a.w, b[c].w, acc[i][c]);
      }
    }
  }
}

// dynamic smem: xs[32][P] | Ws[128][P] (reused as ys[32][P] per 128-column block of the output)
__global__ void __launch_bounds__(kCondThreads)
cond_fwd_kernel(const float* __restrict__ params, long long S, int Zin, int Zout, const int4* __restrict__ tiles,
                const int* __restrict__ rows, const float* __restrict__ x, int ldx, float* __restrict__ out,
                __nv_bfloat16* __restrict__ out16, float* __restrict__ pre, int ldo, const int* __restrict__ ooff,
                float* __restrict__ rstd, int B, int layer_norm, int relu) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;
  float* Ws = smem + kCondRows * kCondP;
  float* ys = Ws + kCondThreads * kCondP;            // [32][Zout + 4]: the whole output row block for LayerNorm
  __shared__ int row_s[kCondRows];
  pdl_sync();
  const int4 t = tiles[blockIdx.x];
  const int slot = t.x, start = t.y, count = t.z, cond = t.w & 0xFFFF;
  const float* W = params + (long long)slot * S;
  const float* bias = W + (long long)Zout * Zin;
  const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
  const int PY = Zout + 4;
  if (tid < kCondRows) row_s[tid] = tid < count ? rows[start + tid] : -1;
  const bool active = 8 * ty < count;
  for (int j0 = 0; j0 < Zout; j0 += kCondThreads) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    for (int k0 = 0; k0 < Zin; k0 += kCondK) {
      __syncthreads();
      cond_stage_x(xs - k0, x, ldx, row_s, k0, Zin, ty, tx);      // (chunk k0 lands at columns 0..127)
      cond_stage_rows(Ws, W + (long long)j0 * Zin, Zin, min(kCondThreads, Zout - j0), k0, Zin, ty, tx);
      __syncthreads();
      if (active) cond_mma_nt(acc, xs, Ws, ty, tx);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx + 32 * c;
      if (j < Zout) {
        const float bj = bias[j];
#pragma unroll
        for (int i = 0; i < 8; ++i) ys[(8 * ty + i) * PY + j] = acc[i][c] + bj;
      }
    }
  }
  __syncthreads();
  // LayerNorm (no affine, eps 1e-5, biased variance) + activation: one warp per row
  const int col0 = ooff[cond];
  for (int r = ty; r < count; r += kCondThreads / 32) {
    const int row = row_s[r];
    const float* y = ys + r * PY;
    float mean = 0.f, rs = 1.f;
    if (layer_norm) {
      float s = 0.f;
      for (int j = tx; j < Zout; j += 32) s += y[j];
      s = warp_sum(s);
      mean = s / (float)Zout;
      float q = 0.f;
      for (int j = tx; j < Zout; j += 32) {
        const float d = y[j] - mean;
        q = fmaf(d, d, q);
      }
      q = warp_sum(q);
      rs = rsqrtf(q / (float)Zout + 1e-5f);
      if (tx == 0) rstd[(long long)cond * B + row] = rs;
    }
    for (int j = tx; j < Zout; j += 32) {
      const float h = (y[j] - mean) * rs;
      const float o = relu ? fmaxf(h, 0.f) : h;
      const long long at = (long long)row * ldo + col0 + j;
      pre[at] = h;
      out[at] = o;
      if (out16) out16[at] = __float2bfloat16(o);
    }
  }
}

// dynamic smem: xs[32][Zin + 4] | dys[32][Zout + 4] | Ws[128][P]
// (Zin, Zout <= 128 per pass of the two products; wider blocks loop over 128-wide pieces)
__global__ void __launch_bounds__(kCondThreads)
cond_bwd_kernel(const float* __restrict__ params, float* __restrict__ grads, long long S, int Zin, int Zout,
                const int4* __restrict__ tiles, const int* __restrict__ rows, const float* __restrict__ x, int ldx,
                const float* __restrict__ dout, const float* __restrict__ pre, int ldo,
                const int* __restrict__ ooff, const float* __restrict__ rstd, int B, float* __restrict__ dx,
                int lddx, const int* __restrict__ dxoff, int layer_norm, int relu) {
  extern __shared__ __align__(16) float smem[];
  const int PX = Zin + 4, PY = Zout + 4;
  float* xs = smem;                           // [32][PX]   gathered inputs
  float* dys = xs + kCondRows * PX;           // [32][PY]   dy
  float* Ws = dys + kCondRows * PY;           // [128][kCondP]  transposed W piece
  __shared__ int row_s[kCondRows];
  pdl_sync();
  const int4 t = tiles[blockIdx.x];
  const int slot = t.x, start = t.y, count = t.z, cond = t.w & 0xFFFF;
  const bool sole = (t.w >> 16) != 0;     // the only tile of its slot: plain stores instead of atomic adds
  const float* W = params + (long long)slot * S;
  float* gW = grads + (long long)slot * S;
  float* gb = gW + (long long)Zout * Zin;
  const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
  if (tid < kCondRows) row_s[tid] = tid < count ? rows[start + tid] : -1;
  __syncthreads();
  for (int k0 = 0; k0 < Zin; k0 += kCondK) cond_stage_x(xs, x, ldx, row_s, k0, Zin, ty, tx, PX, false);
  // dy of the tile's rows (LayerNorm backward without affine: dy = rstd (g - mean(g) - h mean(g h)))
  const int col0 = ooff[cond];
  if (Zout <= 128) {
    // a warp's 8 rows at once: all loads of the rows' LayerNorm outputs and output gradients are in flight together
    float h[8][4], g[8][4], rs[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int row = row_s[ty + 4 * u];
      rs[u] = (row >= 0 && layer_norm) ? rstd[(long long)cond * B + row] : 1.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = tx + 32 * c;
        const bool ok = row >= 0 && j < Zout;
        const long long at = (long long)(row >= 0 ? row : 0) * ldo + col0 + j;
        h[u][c] = ok ? pre[at] : 0.f;
        g[u][c] = ok ? dout[at] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (relu && h[u][c] <= 0.f) g[u][c] = 0.f;
        m1 += g[u][c];
        m2 = fmaf(g[u][c], h[u][c], m2);
      }
      if (layer_norm) {
        m1 = warp_sum(m1) / (float)Zout;
        m2 = warp_sum(m2) / (float)Zout;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = tx + 32 * c;
        if (j < Zout) dys[(ty + 4 * u) * PY + j] = layer_norm ? rs[u] * (g[u][c] - m1 - h[u][c] * m2) : g[u][c];
      }
    }
  } else {
    for (int r = ty; r < kCondRows; r += kCondThreads / 32) {
      const int row = row_s[r];
      if (row < 0) {
        for (int j = tx; j < Zout; j += 32) dys[r * PY + j] = 0.f;
        continue;
      }
      const long long at = (long long)row * ldo + col0;
      float m1 = 0.f, m2 = 0.f, rs = 1.f;
      if (layer_norm) {
        for (int j = tx; j < Zout; j += 32) {
          const float h = pre[at + j];
          const float g = (relu && h <= 0.f) ? 0.f : dout[at + j];
          m1 += g;
          m2 = fmaf(g, h, m2);
        }
        m1 = warp_sum(m1) / (float)Zout;
        m2 = warp_sum(m2) / (float)Zout;
        rs = rstd[(long long)cond * B + row];
      }
      for (int j = tx; j < Zout; j += 32) {
        const float h = pre[at + j];
        const float g = (relu && h <= 0.f) ? 0.f : dout[at + j];
        dys[r * PY + j] = layer_norm ? rs * (g - m1 - h * m2) : g;
      }
    }
  }
  __syncthreads();
  // db_s[j] (+)= sum_r dy[r][j]
  for (int j = tid; j < Zout; j += kCondThreads) {
    float s = 0.f;
    for (int r = 0; r < count; ++r) s += dys[r * PY + j];
    if (sole) gb[j] = s; else atomicAdd(gb + j, s);
  }
  // dW_s[j][k] (+)= sum_r dy[r][j] x[r][k]: thread = columns k in {tx + 32 c}, 8 rows j per pass; the loop over r
  // stops at the tile's row count
  for (int j0 = 8 * ty; j0 < Zout; j0 += 8 * (kCondThreads / 32)) {
    for (int kb = 0; kb < Zin; kb += 128) {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
      for (int r = 0; r < count; ++r) {
        float xv[4], dv[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = kb + tx + 32 * c;
          xv[c] = k < Zin ? xs[r * PX + k] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) dv[i] = (j0 + i) < Zout ? dys[r * PY + j0 + i] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(dv[i], xv[c], acc[i][c]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = j0 + i, k = kb + tx + 32 * c;
          if (j < Zout && k < Zin) {
            if (sole) gW[(long long)j * Zin + k] = acc[i][c]; else atomicAdd(gW + (long long)j * Zin + k, acc[i][c]);
          }
        }
    }
  }
  // dx[r][k] = sum_j dy[r][j] W[j][k]: W^T pieces [128 k][128 j] staged as Ws[k][j] (a transposing copy: the global
  // rows j are read coalesced, 4 k per lane, and scattered to 4 smem rows)
  const int xoff = dxoff[cond];
  const bool active = 8 * ty < count;
  for (int kb = 0; kb < Zin; kb += kCondThreads) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    for (int jb = 0; jb < Zout; jb += kCondK) {
      __syncthreads();
      for (int jj0 = ty; jj0 < kCondK; jj0 += 8 * (kCondThreads / 32)) {
        float w[8][4];       // 32 independent loads in flight per thread before the first store
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = jb + jj0 + u * (kCondThreads / 32);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int k = kb + tx + 32 * c;   // coalesced along k
            w[u][c] = (j < Zout && k < Zin) ? __ldg(W + (long long)j * Zin + k) : 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int c = 0; c < 4; ++c) Ws[(tx + 32 * c) * kCondP + jj0 + u * (kCondThreads / 32)] = w[u][c];
      }
      __syncthreads();
      if (active) {
        // acc[i][c] += sum_jj dys[8 ty + i][jb + jj] * Ws[tx + 32 c][jj]
#pragma unroll 2
        for (int jj = 0; jj < kCondK; jj += 4) {
          if (jb + jj >= Zout) break;
          float4 b[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(Ws + (tx + 32 * c) * kCondP + jj);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float* dp = dys + (8 * ty + i) * PY + jb + jj;
            const float a0 = dp[0], a1 = (jb + jj + 1 < Zout) ? dp[1] : 0.f, a2 = (jb + jj + 2 < Zout) ? dp[2] : 0.f,
                        a3 = (jb + jj + 3 < Zout) ? dp[3] : 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              acc[i][c] = fmaf(a0, b[c].x, acc[i][c]);
              acc[i][c] = fmaf(a1, b[c].y, acc[i][c]);
              acc[i][c] = fmaf(a2, b[c].z, acc[i][c]);
              acc[i][c] = fmaf(a3, b[c].w, acc[i][c]);
            }
          }
        }
      }
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = row_s[8 * ty + i];
        if (row < 0) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = kb + tx + 32 * c;
          if (k < Zin) dx[(long long)row * lddx + xoff + k] = acc[i][c];
        }
      }
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
  __shared__ float bc[2];
  pdl_sync();
  const int slot = present[blockIdx.x];
  if (threadIdx.x == 0) {
    const double t = (double)(steps[slot] + 1);
    bc[0] = (float)(1.0 - pow(b1, t));
    bc[1] = (float)(1.0 / sqrt(1.0 - pow(b2, t)));
  }
  __syncthreads();
  const float bc1 = bc[0], isb2 = bc[1];
  float coef = 1.f;
  if (max_norm > 0.f && norm_sq) {
    const float total = (float)sqrt(*norm_sq) * fabsf(grad_scale);
    coef = fminf(1.f, max_norm / (total + 1e-6f));
  }
  const float gs = coef * grad_scale;
  const float fb1 = (float)b1, fb2 = (float)b2;
  const long long base = (long long)slot * S;
  float4* p4 = reinterpret_cast<float4*>(params + base);
  float4* m4 = reinterpret_cast<float4*>(m + base);
  float4* v4 = reinterpret_cast<float4*>(v + base);
  const float4* g4 = reinterpret_cast<const float4*>(grads + base);
  for (long long i = blockIdx.y * (long long)blockDim.x + threadIdx.x; i < S / 4; i += (long long)gridDim.y * blockDim.x) {
    float4 p = p4[i], mm = m4[i], vv = v4[i];
    const float4 g = g4[i];
    adam_one(p.x, g.x, mm.x, vv.x, gs, lr, fb1, fb2, eps, wd, bc1, isb2);
    adam_one(p.y, g.y, mm.y, vv.y, gs, lr, fb1, fb2, eps, wd, bc1, isb2);
    adam_one(p.z, g.z, mm.z, vv.z, gs, lr, fb1, fb2, eps, wd, bc1, isb2);
    adam_one(p.w, g.w, mm.w, vv.w, gs, lr, fb1, fb2, eps, wd, bc1, isb2);
    p4[i] = p; m4[i] = mm; v4[i] = vv;
  }
}

__global__ void cond_step_inc_kernel(int* __restrict__ steps, const int* __restrict__ present, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) steps[present[i]] += 1;
}
}  // namespace cmmvae

using namespace cmmvae;

static size_t cond_smem_fwd(int Zin, int Zout) {
  return (size_t)(kCondRows * kCondP + kCondThreads * kCondP + kCondRows * (Zout + 4)) * sizeof(float);
}
static size_t cond_smem_bwd(int Zin, int Zout) {
  return (size_t)(kCondRows * (Zin + 4) + kCondRows * (Zout + 4) + kCondThreads * kCondP) * sizeof(float);
}

extern "C" int cmmvae_cond_fwd(const float* params, long long slot_stride, int Zin, int Zout, const int32_t* tiles,
                               int n_tiles, const int32_t* rows, const float* x, int ldx, float* out, void* out_bf16,
                               float* pre, int ldo, const int32_t* out_col, float* rstd, int B, int layer_norm, int relu,
                               void* stream) {
  if (n_tiles <= 0) return 0;
  CMMVAE_REQUIRE(params && tiles && rows && x && out && pre && out_col && rstd, "cond_fwd: null pointer");
  CMMVAE_REQUIRE(Zin > 0 && Zout > 0 && Zin % 4 == 0 && Zout % 4 == 0 && slot_stride % 4 == 0 &&
                     slot_stride >= (long long)Zout * Zin + Zout,
                 "cond_fwd: bad sizes (block widths must be multiples of 4)");
  const size_t smem = cond_smem_fwd(Zin, Zout);
  CMMVAE_REQUIRE(smem <= 200 * 1024, "cond_fwd: block too wide for the shared-memory tile");
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
  CMMVAE_REQUIRE(Zin > 0 && Zout > 0 && Zin % 4 == 0 && Zout % 4 == 0 && slot_stride % 4 == 0,
                 "cond_bwd: bad sizes (block widths must be multiples of 4)");
  const size_t smem = cond_smem_bwd(Zin, Zout);
  CMMVAE_REQUIRE(smem <= 200 * 1024, "cond_bwd: block too wide for the shared-memory tile");
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
  const int chunks = (int)((slot_stride / 4 + 255) / 256);
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
