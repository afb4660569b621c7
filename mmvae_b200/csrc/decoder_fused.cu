// K5+K6+K7 fused: expert-decoder output GEMM (tcgen05, accumulators in TMEM) whose epilogue adds the
// bias, applies ReLU, reduces the sum-MSE against the CSR batch and emits dlogits (bf16).  The
// [cells x genes] reconstruction never reaches HBM.
//
//   logits[b,g] = h[b,:] . Wout[g,:] + bout[g]      xhat = relu(logits)
//   loss       += sum_g xhat^2  +  sum_{nz} (x^2 - 2 x xhat)              (= sum (xhat - x)^2)
//   dlogits     = 2 (xhat - x) 1[logits > 0]
//
// Persistent kernel, one CTA per SM, 128 x 256 output tiles, static round-robin tile order with the
// cell-block index fastest (so CTAs running concurrently share the same Wout tiles in L2).
//   warp 0      TMA producer (3-stage ring of 128x64 h tiles + 256x64 Wout tiles, 128B swizzle)
//   warp 1      tcgen05.mma issuer; two 256-column TMEM accumulators ping-pong with the epilogue
//   warps 2..5  epilogue: thread = cell.  The cell's CSR entries inside the gene window are located
//               through a per-(cell, gene-tile) pointer table (built once per batch by
//               tile_ptr_kernel, bit-exact with crow/col) and staged in shared memory while the MMA
//               runs; accumulator chunks go TMEM -> registers -> per-thread smem column, where the
//               sparse entries are patched in, then out as bf16.
#include "tc.cuh"

namespace cmmvae {

using namespace tc;

constexpr int DBM = 128, DBN = 256, DBK = 64, DSTAGES = 3;
constexpr int DCAP = 48;  // CSR entries per (cell, gene tile) staged in smem; the rest stream from global
constexpr int kDecThreads = 192;

struct DecSmem {
  static constexpr int kABytes = DBM * DBK * 2;           // 16 KB
  static constexpr int kBBytes = DBN * DBK * 2;           // 32 KB
  static constexpr int kStageBytes = kABytes + kBBytes;   // 48 KB
  static constexpr int kStageOff = 0;
  static constexpr int kStagingOff = DSTAGES * kStageBytes;          // float [32][128]
  static constexpr int kEntValOff = kStagingOff + 32 * 128 * 4;      // float [DCAP][128]
  static constexpr int kEntColOff = kEntValOff + DCAP * 128 * 4;     // uint8 [DCAP][128]
  static constexpr int kBiasOff = kEntColOff + DCAP * 128;           // float [256]
  static constexpr int kBarOff = kBiasOff + DBN * 4;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};

struct DecParams {
  int B, G, H;
  const float* bout;
  const int32_t* crow;
  const int32_t* col;
  const float* val;
  const int32_t* tile_ptr;  // [B][num_n + 1]
  __nv_bfloat16* dl;
  int ldd;
  double* loss_sum;
  int num_m, num_n;
};

// tile_ptr[b][t] = first CSR position of row b whose column is >= t * DBN   (t = 0..num_n)
__global__ void tile_ptr_kernel(const int32_t* __restrict__ crow, const int32_t* __restrict__ col, int B, int num_n,
                                int32_t* __restrict__ tp) {
  const long long n = (long long)B * (num_n + 1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / (num_n + 1)), t = (int)(i % (num_n + 1));
    int lo = crow[b], hi = crow[b + 1];
    const int key = t * DBN;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (col[mid] < key) lo = mid + 1; else hi = mid;
    }
    tp[i] = lo;
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kDecThreads, 1)
decoder_mse_fused_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW,
                         const DecParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzled tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* staging = reinterpret_cast<float*>(smem + DecSmem::kStagingOff);
  float* ent_val = reinterpret_cast<float*>(smem + DecSmem::kEntValOff);
  uint8_t* ent_col = smem + DecSmem::kEntColOff;
  float* s_bias = reinterpret_cast<float*>(smem + DecSmem::kBiasOff);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + DecSmem::kBarOff);
  uint64_t* empty_bar = full_bar + DSTAGES;
  uint64_t* tmem_full = empty_bar + DSTAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.H + DBK - 1) / DBK;
  const int num_tiles = p.num_m * p.num_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmH);
    prefetch_tmap(&tmW);
    for (int s = 0; s < DSTAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t % p.num_m) * DBM, n0 = (t / p.num_m) * DBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * DecSmem::kStageBytes;
          mbar_expect_tx(&full_bar[stage], DecSmem::kStageBytes);
          tma_load_2d(sA, &tmH, &full_bar[stage], kb * DBK, m0);
          tma_load_2d(sA + DecSmem::kABytes, &tmW, &full_bar[stage], kb * DBK, n0);
          if (++stage == DSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(DBM, DBN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * DBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * DecSmem::kStageBytes);
          const uint32_t sB = sA + DecSmem::kABytes;
#pragma unroll
          for (int k = 0; k < DBK / 16; ++k)
            umma_bf16(tmem_d, make_desc_sw128(sA + k * 32, 16, 1024), make_desc_sw128(sB + k * 32, 16, 1024), idesc,
                      (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == DSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;
    const int row = q * 32 + lane;           // TMEM lane == row inside the tile
    const int et = threadIdx.x - 64;         // 0..127 among epilogue threads
    double loss_acc = 0.0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_blk = t % p.num_m, n_blk = t / p.num_m;
      const int m0 = m_blk * DBM, n0 = n_blk * DBN;
      const int b = m0 + row;
      const bool row_ok = b < p.B;

      // bias tile -> smem (zero beyond G so padded columns give xhat = 0)
      named_bar_sync(1, 128);
      for (int j = et; j < DBN; j += 128) s_bias[j] = (n0 + j < p.G) ? __ldg(p.bout + n0 + j) : 0.f;
      // this cell's CSR entries inside [n0, n0 + DBN): stage up to DCAP of them while the MMA runs
      int p0 = 0, p1 = 0;
      if (row_ok) {
        const int32_t* tp = p.tile_ptr + (size_t)b * (p.num_n + 1) + n_blk;
        p0 = __ldg(tp);
        p1 = __ldg(tp + 1);
      }
      const int cnt = p1 - p0;
      const int staged = min(cnt, DCAP);
#pragma unroll 4
      for (int k = 0; k < staged; ++k) {
        ent_col[k * 128 + row] = (uint8_t)(__ldg(p.col + p0 + k) - n0);
        ent_val[k * 128 + row] = __ldg(p.val + p0 + k);
      }
      named_bar_sync(1, 128);

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      float part = 0.f;
      int ek = 0;  // next entry of this row
#pragma unroll 1
      for (int c = 0; c < DBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * DBN + c * 32), r);
        tmem_ld_wait();
        // dense part: xhat = relu(acc + bias); staged column-major per thread (conflict free)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float xh = fmaxf(__uint_as_float(r[j]) + s_bias[c * 32 + j], 0.f);
          part = fmaf(xh, xh, part);
          staging[j * 128 + row] = xh;
        }
        // sparse part: patch the entries of this 32-gene chunk
        const int chunk_end = (c + 1) * 32;
        while (ek < cnt) {
          int cj;
          float x;
          if (ek < DCAP) {
            cj = ent_col[ek * 128 + row];
            x = ent_val[ek * 128 + row];
          } else {
            cj = __ldg(p.col + p0 + ek) - n0;
            x = __ldg(p.val + p0 + ek);
          }
          if (cj >= chunk_end) break;
          const int j = cj - c * 32;
          const float xh = staging[j * 128 + row];
          part += x * x - 2.f * x * xh;
          staging[j * 128 + row] = xh > 0.f ? xh - x : 0.f;
          ++ek;
        }
        // out: dlogits = 2 * staged, bf16, 8 columns (16 bytes) per store
        if (row_ok) {
          __nv_bfloat16* drow = p.dl + (size_t)b * p.ldd + n0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (n0 + c * 32 + j + 8 <= p.ldd) {
              uint4 o;
              o.x = pack_bf16(2.f * staging[(j + 0) * 128 + row], 2.f * staging[(j + 1) * 128 + row]);
              o.y = pack_bf16(2.f * staging[(j + 2) * 128 + row], 2.f * staging[(j + 3) * 128 + row]);
              o.z = pack_bf16(2.f * staging[(j + 4) * 128 + row], 2.f * staging[(j + 5) * 128 + row]);
              o.w = pack_bf16(2.f * staging[(j + 6) * 128 + row], 2.f * staging[(j + 7) * 128 + row]);
              *reinterpret_cast<uint4*>(drow + j) = o;
            }
          }
        }
      }
      if (row_ok) loss_acc += (double)part;
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    // one atomic per warp
    loss_acc = warp_sum(loss_acc);
    if (lane == 0 && loss_acc != 0.0) atomicAdd(p.loss_sum, loss_acc);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                   uint32_t box_outer);

}  // namespace cmmvae

using namespace cmmvae;

extern "C" size_t cmmvae_decoder_mse_fused_workspace_bytes(int B, int G) {
  const int num_n = (G + DBN - 1) / DBN;
  return sizeof(int32_t) * (size_t)B * (size_t)(num_n + 1);
}

extern "C" int cmmvae_decoder_mse_fused(const void* h, int ldh, const void* Wout, int ldw, const float* bout, int B,
                                        int G, int H, const int32_t* crow, const int32_t* col, const float* val,
                                        void* dlogits_bf16, int ldd, double* loss_sum, void* workspace,
                                        void* stream) {
  CMMVAE_REQUIRE(B > 0 && G > 0 && H > 0, "decoder_mse_fused: bad shape");
  CMMVAE_REQUIRE(ldh % 8 == 0 && ldw % 8 == 0 && ldd % 8 == 0 && ldd >= G,
                 "decoder_mse_fused: ldh/ldw/ldd must be multiples of 8 and ldd >= G");
  CMMVAE_REQUIRE((((uintptr_t)h | (uintptr_t)Wout | (uintptr_t)dlogits_bf16) & 15) == 0,
                 "decoder_mse_fused: pointers must be 16-byte aligned");
  CMMVAE_REQUIRE(workspace && loss_sum, "decoder_mse_fused: workspace/loss_sum missing");
  cudaStream_t st = (cudaStream_t)stream;
  DecParams p;
  p.B = B; p.G = G; p.H = H; p.bout = bout; p.crow = crow; p.col = col; p.val = val;
  p.tile_ptr = (const int32_t*)workspace;
  p.dl = (__nv_bfloat16*)dlogits_bf16; p.ldd = ldd; p.loss_sum = loss_sum;
  p.num_m = (B + DBM - 1) / DBM;
  p.num_n = (G + DBN - 1) / DBN;
  CUtensorMap tmH, tmW;
  if (int rc = make_tmap_bf16(&tmH, h, (uint64_t)H, (uint64_t)B, (uint64_t)ldh, DBK, DBM)) return rc;
  if (int rc = make_tmap_bf16(&tmW, Wout, (uint64_t)H, (uint64_t)G, (uint64_t)ldw, DBK, DBN)) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(decoder_mse_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         DecSmem::kTotal);
    if (e != cudaSuccess) {
      set_error("decoder_mse_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  cudaMemsetAsync(loss_sum, 0, sizeof(double), st);
  {
    const long long n = (long long)B * (p.num_n + 1);
    long long want = (n + 255) / 256;
    int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
    tile_ptr_kernel<<<blocks, 256, 0, st>>>(crow, col, B, p.num_n, (int32_t*)workspace);
    if (int rc = check_launch("tile_ptr")) return rc;
  }
  const int num_tiles = p.num_m * p.num_n;
  const int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
  decoder_mse_fused_kernel<<<grid, kDecThreads, DecSmem::kTotal, st>>>(tmH, tmW, p);
  return check_launch("decoder_mse_fused");
}
