// K5+K6+K7 fused: expert-decoder output GEMM (tcgen05, accumulators in TMEM) whose epilogue adds the
// bias, applies ReLU, reduces the sum-MSE against the CSR batch and emits dlogits (bf16).  The
// [cells x genes] reconstruction never reaches HBM.
//
//   logits[b,g] = h[b,:] . Wout[g,:] + bout[g]      xhat = relu(logits)
//   loss       += sum_g xhat^2  +  sum_{nz} (x^2 - 2 x xhat)              (= sum (xhat - x)^2)
//   dlogits     = 2 (xhat - x) 1[logits > 0]
//
// Persistent kernel, one CTA per SM, 128 x 256 output tiles, static round-robin tile order with the
// cell-block index fastest (so CTAs running concurrently share the same Wout tiles in L2).  576 threads:
//   warp 16      TMA producer (3-stage ring of 128x64 h tiles + 256x64 Wout tiles, 128B swizzle)
//   warp 17      tcgen05.mma issuer (highest warp id = issue priority); two 256-column TMEM accumulators
//                ping-pong with the epilogue
//   warps 0..15  epilogue, 16 warps so that draining a tile takes less time than computing the next one:
//                thread = (cell, 64-gene window).  The cell's CSR entries of the window are found through
//                the per-(cell, 64-gene window) pointer table (window-major, shared with the tensor-pipe
//                SpMM; bit-exact with crow/col) and prefetched into registers while the MMA runs;
//                accumulator sub-chunks go TMEM -> registers -> a per-thread shared-memory column where
//                the sparse entries are patched in, then as bf16 into a 128B-swizzled [128 x 64] output tile
//                per window that one thread hands to a TMA bulk store (full-line writes, edges clipped by
//                the tensor map) -- per-thread 32-byte global stores cost 0.05 ms of L1 wavefronts here.
#include "tc.cuh"

namespace cmmvae {

using namespace tc;

constexpr int DBM = 128, DBN = 256, DBK = 64, DSTAGES = 3;
constexpr int kDecEpiWarps = 16;
constexpr int kDecThreads = 64 + 32 * kDecEpiWarps;   // 576
constexpr int DE = 8;   // CSR entries per (cell, window) prefetched into registers; the rest stream from global

struct DecSmem {
  static constexpr int kABytes = DBM * DBK * 2;           // 16 KB
  static constexpr int kBBytes = DBN * DBK * 2;           // 32 KB
  static constexpr int kStageBytes = kABytes + kBBytes;   // 48 KB
  static constexpr int kOutOff = DSTAGES * kStageBytes;                     // 4 x [128 rows][128 B] bf16, SW128
  static constexpr int kStagingOff = kOutOff + 4 * DBM * 128;               // float [8][512]
  static constexpr int kBarOff = kStagingOff + 8 * 32 * kDecEpiWarps * 4;   // + 16 KB
  static constexpr int kBiasOff = kBarOff + 256;          // float [2][256]: the tile's output biases, double buffered
  static constexpr int kTotal = kBiasOff + 2 * DBN * 4;
};

struct DecParams {
  int B, G, H;
  const float* bout;
  const int32_t* col;
  const float* val;
  const int32_t* tp;   // [ntp][B] window-major pointer table (64-gene windows)
  int ntp;
  __nv_bfloat16* dl;
  int ldd;
  double* loss_sum;   // [1], or one slot per block of `loss_rows` cells (data parallel: one slot per source rank)
  int loss_rows;      // 0: everything into loss_sum[0]
  int num_m, num_n;
};

__global__ void __launch_bounds__(kDecThreads, 1)
decoder_mse_fused_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW,
                         const __grid_constant__ CUtensorMap tmD, const DecParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzled tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* staging = reinterpret_cast<float*>(smem + DecSmem::kStagingOff);
  float* bias_s = reinterpret_cast<float*>(smem + DecSmem::kBiasOff);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + DecSmem::kBarOff);
  uint64_t* empty_bar = full_bar + DSTAGES;
  uint64_t* tmem_full = empty_bar + DSTAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.H + DBK - 1) / DBK;
  const int num_tiles = p.num_m * p.num_n;

  constexpr int kTmaWarp = kDecEpiWarps, kMmaWarp = kDecEpiWarps + 1;   // MMA issuer = highest warp id (priority)
  if (warp == kTmaWarp && lane == 0) {
    prefetch_tmap(&tmH);
    prefetch_tmap(&tmW);
    for (int s = 0; s < DSTAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kDecEpiWarps);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // prologue overlapped the previous kernel; its results are visible from here on

  if (warp == kTmaWarp) {
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
  } else if (warp == kMmaWarp) {
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
      pdl_trigger();   // every MMA of this CTA is issued: the next kernel in the stream may move in
    }
  } else {
    // ===== epilogue: thread = (cell row, 64-gene window) =====
    const int e = warp;                      // 0..15
    const int q = warp & 3;                  // TMEM lane group this warp may read
    const int w = e >> 2;                    // 64-gene window of the tile: consecutive warps cover all lane groups
    const int row = q * 32 + lane;           // TMEM lane == row inside the tile
    const int et = e * 32 + lane;            // 0..511
    float* st = staging + et;                // element j of this thread's column: st[j * 512]
    uint8_t* out_tile = smem + DecSmem::kOutOff + w * (DBM * 128);   // this window's [128][64] bf16 tile
    const int gt = (e & 3) * 32 + lane;      // 0..127 inside the window's 4-warp group
    constexpr int SS = 32 * kDecEpiWarps;
    double loss_acc = 0.0;
    // CSR entries of (this thread's cell, this thread's window) for a tile, fetched one tile ahead so the
    // two dependent L2 round trips (pointer table, then entries) hide behind the previous tile's drain
    auto fetch = [&](int t, int& p0, int& cnt, int (&ecol)[DE], float (&eval)[DE]) {
      p0 = 0;
      int p1 = 0;
      if (t < num_tiles) {
        const int b = (t % p.num_m) * DBM + row;
        const int g0 = (t / p.num_m) * DBN + w * 64;
        const int win = g0 >> 6;
        if (b < p.B && win < p.ntp - 1) {
          p0 = __ldg(p.tp + (size_t)win * p.B + b);
          p1 = __ldg(p.tp + (size_t)(win + 1) * p.B + b);
        }
        cnt = p1 - p0;
#pragma unroll
        for (int u = 0; u < DE; ++u) {
          const bool ok = u < cnt;
          ecol[u] = ok ? __ldg(p.col + p0 + u) - g0 : 1 << 20;
          eval[u] = ok ? __ldg(p.val + p0 + u) : 0.f;
        }
      } else {
        cnt = 0;
      }
    };
    int p0, cnt, np0, ncnt;
    int ecol[DE], necol[DE];
    float eval[DE], neval[DE];
    fetch(blockIdx.x, p0, cnt, ecol, eval);
    int it = 0;
    int slot = 0;       // loss slot the partial sum in loss_acc belongs to
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_blk = t % p.num_m, n_blk = t / p.num_m;
      const int m0 = m_blk * DBM, n0 = n_blk * DBN;
      const int b = m0 + row;
      const bool row_ok = b < p.B;
      if (p.loss_rows > 0) {
        // slot of this thread's cell.  loss_rows % 32 == 0 (the usual case: cells per rank a multiple of 32): the
        // 32 rows of a warp share the slot, so a slot change flushes one warp-reduced atomic; otherwise per thread
        const int s_new = row_ok ? b / p.loss_rows : slot;
        if (s_new != slot) {
          if ((p.loss_rows & 31) == 0) {
            const double w_sum = warp_sum(loss_acc);
            if (lane == 0 && w_sum != 0.0) atomicAdd(p.loss_sum + slot, w_sum);
          } else if (loss_acc != 0.0) {
            atomicAdd(p.loss_sum + slot, loss_acc);
          }
          loss_acc = 0.0;
          slot = s_new;
        }
      }
      const int g0 = n0 + w * 64;            // first gene of this thread's window
      fetch(t + gridDim.x, np0, ncnt, necol, neval);

      // the window's 64 output biases go to shared memory once per tile (the per-chunk global loads in the drain
      // loop were its largest stall); two buffers: a warp is at most one tile ahead of its window group
      float* bias_w = bias_s + (it & 1) * DBN + w * 64;
      if (gt < 64) bias_w[gt] = (g0 + gt < p.G) ? __ldg(p.bout + g0 + gt) : 0.f;
      // the window's output tile (shared by the 4 warps of this window) must have been read by the previous
      // tile's TMA store before it is overwritten
      if (gt == 0) tma_store_wait_read();
      named_bar_sync(1 + w, 128);
      mbar_wait_relaxed(&tmem_full[acc], acc_phase);
      tc_fence_after();
      float part = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {          // eight 8-column sub-chunks of the window (one 16-byte bf16 chunk each)
        uint32_t r[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * DBN + w * 64 + c * 8), r);
        float bias[8];
        {
          const float4 b0 = *reinterpret_cast<const float4*>(bias_w + c * 8);       // same address for the whole warp
          const float4 b1 = *reinterpret_cast<const float4*>(bias_w + c * 8 + 4);
          bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
          bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
        }
        tmem_ld_wait();
        // dense part: xhat = relu(acc + bias), staged in this thread's smem column (conflict free)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = fmaxf(__uint_as_float(r[j]) + bias[j], 0.f);
          part = fmaf(xh, xh, part);
          st[j * SS] = xh;
        }
        // sparse part: patch the entries of this 8-gene sub-chunk
        const int lo = c * 8;
#pragma unroll
        for (int u = 0; u < DE; ++u) {
          const int j = ecol[u] - lo;
          if (j >= 0 && j < 8) {
            const float x = eval[u];
            const float xh = st[j * SS];
            part += x * x - 2.f * x * xh;
            st[j * SS] = xh > 0.f ? xh - x : 0.f;
          }
        }
        for (int k = DE; k < cnt; ++k) {     // windows with more than DE entries (dense batches)
          const int j = __ldg(p.col + p0 + k) - g0 - lo;
          if (j >= 0 && j < 8) {
            const float x = __ldg(p.val + p0 + k);
            const float xh = st[j * SS];
            part += x * x - 2.f * x * xh;
            st[j * SS] = xh > 0.f ? xh - x : 0.f;
          }
        }
        // out: dlogits = 2 * staged as bf16 -> chunk c of this row in the 128B-swizzled output tile
        uint4 o;
        o.x = pack_bf16(2.f * st[0 * SS], 2.f * st[1 * SS]);
        o.y = pack_bf16(2.f * st[2 * SS], 2.f * st[3 * SS]);
        o.z = pack_bf16(2.f * st[4 * SS], 2.f * st[5 * SS]);
        o.w = pack_bf16(2.f * st[6 * SS], 2.f * st[7 * SS]);
        *reinterpret_cast<uint4*>(out_tile + row * 128 + ((c ^ (row & 7)) << 4)) = o;
      }
      // publish the window: generic-proxy writes -> async proxy, then one thread stores the 128 x 64 box
      // (rows beyond B and columns beyond ldd are clipped by the tensor map)
      fence_proxy_async();
      named_bar_sync(1 + w, 128);
      if (gt == 0) {
        tma_store_2d(&tmD, out_tile, g0, m0);
        tma_store_commit();
      }
      if (row_ok) loss_acc += (double)part;
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      p0 = np0; cnt = ncnt;
#pragma unroll
      for (int u = 0; u < DE; ++u) { ecol[u] = necol[u]; eval[u] = neval[u]; }
    }
    if (gt == 0) tma_store_wait_all();       // smem must outlive the last bulk store
    if (p.loss_rows > 0 && (p.loss_rows & 31) != 0) {
      if (loss_acc != 0.0) atomicAdd(p.loss_sum + slot, loss_acc);
    } else if (p.loss_rows > 0) {
      loss_acc = warp_sum(loss_acc);
      if (lane == 0 && loss_acc != 0.0) atomicAdd(p.loss_sum + slot, loss_acc);
    } else {
      // one atomic per warp
      loss_acc = warp_sum(loss_acc);
      if (lane == 0 && loss_acc != 0.0) atomicAdd(p.loss_sum, loss_acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                   uint32_t box_outer);
__global__ void tile_ptr64_kernel(const int32_t* __restrict__ rbeg, const int32_t* __restrict__ rend,
                                  const int32_t* __restrict__ col, int B, int ntp, int32_t* __restrict__ tp);

}  // namespace cmmvae

using namespace cmmvae;

extern "C" size_t cmmvae_decoder_mse_fused_workspace_bytes(int B, int G) {
  return sizeof(int32_t) * (size_t)B * (size_t)((G + 63) / 64 + 1);
}

static int decoder_mse(const void* h, int ldh, const void* Wout, int ldw, const float* bout, int B, int G, int H,
                       const int32_t* crow, const int32_t* col, const float* val, const int32_t* tile_ptr,
                       void* dlogits_bf16, int ldd, double* loss_sum, int loss_rows, void* workspace, void* stream) {
  CMMVAE_REQUIRE(B > 0 && G > 0 && H > 0, "decoder_mse_fused: bad shape");
  CMMVAE_REQUIRE(ldh % 8 == 0 && ldw % 8 == 0 && ldd % 8 == 0 && ldd >= G,
                 "decoder_mse_fused: ldh/ldw/ldd must be multiples of 8 and ldd >= G");
  CMMVAE_REQUIRE((((uintptr_t)h | (uintptr_t)Wout | (uintptr_t)dlogits_bf16) & 15) == 0,
                 "decoder_mse_fused: pointers must be 16-byte aligned");
  CMMVAE_REQUIRE(((uintptr_t)bout & 7) == 0, "decoder_mse_fused: bout must be 8-byte aligned");
  CMMVAE_REQUIRE((tile_ptr || workspace) && loss_sum, "decoder_mse_fused: tile_ptr/workspace or loss_sum missing");
  cudaStream_t st = (cudaStream_t)stream;
  DecParams p;
  p.B = B; p.G = G; p.H = H; p.bout = bout; p.col = col; p.val = val;
  p.ntp = (G + 63) / 64 + 1;
  p.dl = (__nv_bfloat16*)dlogits_bf16; p.ldd = ldd; p.loss_sum = loss_sum; p.loss_rows = loss_rows;
  p.num_m = (B + DBM - 1) / DBM;
  p.num_n = (G + DBN - 1) / DBN;
  CUtensorMap tmH, tmW, tmD;
  if (int rc = make_tmap_bf16(&tmH, h, (uint64_t)H, (uint64_t)B, (uint64_t)ldh, DBK, DBM)) return rc;
  if (int rc = make_tmap_bf16(&tmW, Wout, (uint64_t)H, (uint64_t)G, (uint64_t)ldw, DBK, DBN)) return rc;
  if (int rc = make_tmap_bf16(&tmD, dlogits_bf16, (uint64_t)ldd, (uint64_t)B, (uint64_t)ldd, 64, DBM)) return rc;
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
  cudaMemsetAsync(loss_sum, 0, sizeof(double) * (loss_rows > 0 ? (size_t)((B + loss_rows - 1) / loss_rows) : 1), st);
  if (!tile_ptr) {   // build the 64-gene-window pointer table (the tensor-pipe SpMM shares it when it ran)
    int blocks = B < 148 * 16 ? B : 148 * 16;   // one CTA per row
    launch_pdl(tile_ptr64_kernel, dim3(blocks), dim3(256), 0, st, crow, crow + 1, col, B, p.ntp, (int32_t*)workspace);
    if (int rc = check_launch("tile_ptr64")) return rc;
    tile_ptr = (const int32_t*)workspace;
  }
  p.tp = tile_ptr;
  const int num_tiles = p.num_m * p.num_n;
  const int grid = num_tiles < sm_budget() ? num_tiles : sm_budget();
  launch_pdl(decoder_mse_fused_kernel, dim3(grid), dim3(kDecThreads), DecSmem::kTotal, st, tmH, tmW, tmD, p);
  return check_launch("decoder_mse_fused");
}

extern "C" int cmmvae_decoder_mse_fused(const void* h, int ldh, const void* Wout, int ldw, const float* bout, int B,
                                        int G, int H, const int32_t* crow, const int32_t* col, const float* val,
                                        const int32_t* tile_ptr, void* dlogits_bf16, int ldd, double* loss_sum,
                                        void* workspace, void* stream) {
  return decoder_mse(h, ldh, Wout, ldw, bout, B, G, H, crow, col, val, tile_ptr, dlogits_bf16, ldd, loss_sum, 0,
                     workspace, stream);
}

extern "C" int cmmvae_decoder_mse_fused_blocks(const void* h, int ldh, const void* Wout, int ldw, const float* bout,
                                               int B, int G, int H, const int32_t* crow, const int32_t* col,
                                               const float* val, const int32_t* tile_ptr, void* dlogits_bf16, int ldd,
                                               double* loss_sums, int loss_rows, void* workspace, void* stream) {
  CMMVAE_REQUIRE(loss_rows > 0, "decoder_mse_fused_blocks: loss_rows must be positive");
  return decoder_mse(h, ldh, Wout, ldw, bout, B, G, H, crow, col, val, tile_ptr, dlogits_bf16, ldd, loss_sums,
                     loss_rows, workspace, stream);
}
