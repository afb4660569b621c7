// tcgen05 / TMEM / TMA GEMM for the dense layers of the CMMVAE step (K4, K12) on sm_100a.
//
//   C[M,N] = act( opA(A) * opB(B) + bias ) (+ C)        bf16 operands, fp32 accumulation in TMEM
//
// Persistent CTAs walk 128 x BN output tiles (optionally split along K).  Warp roles (192 threads):
//   warps 0..3  epilogue: tcgen05.ld the accumulator (lane = output row), bias / ReLU / accumulate,
//               vectorised stores of f32 and/or bf16 (drains accumulator i while tile i+1 is multiplied)
//   warp 4      TMA producer: streams 128x64 (A) and BNx64 (B) bf16 tiles, 128B-swizzled, through a
//               kStages-deep shared-memory ring guarded by full/empty mbarriers
//   warp 5      allocates TMEM (two accumulators), then one lane issues tcgen05.mma (UMMA 128 x BN x 16) and
//               releases ring slots with tcgen05.commit; it is the highest warp id on purpose (issue priority)
// Operands may be K-major ("row = M or N index, K contiguous") or MN-major (stored transposed):
// both are native UMMA layouts, so backward GEMMs (dX = dY W, dW = dY^T X) need no transposed copies.
#include "tc.cuh"

namespace cmmvae {

using namespace tc;

constexpr int BM = 128;   // UMMA M (cta_group::1)
constexpr int BK = 64;    // 64 bf16 = 128 bytes = one swizzle span
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN <= 128) ? 6 : 4;
  static constexpr int kOutOff = kStages * kStageBytes;        // 2 x [128 rows][128 B] f32 chunks, SW128
  static constexpr int kOutBytes = 2 * BM * 128;
  static constexpr int kBarOff = kOutOff + kOutBytes;
  static constexpr int kTotal = kBarOff + 256;
};

struct GemmParams {
  int M, N, K;
  const float* bias;
  int relu, accumulate;
  float* C32;
  __nv_bfloat16* C16;
  int ldc;
  int vec_ok;  // 16-byte aligned rows for both outputs
  int splits;  // split-K factor (gridDim.z); > 1: partial products are atomically added into a zeroed C32
  double* sumsq;  // optional: += sum of squares of the stored C (gradient-norm fused into the dW GEMM)
  int tma_out;    // f32 output leaves through 128B-swizzled smem chunks + TMA bulk stores (full-line writes)
  // data parallel: output row gm belongs to rank gm / route_rows and is stored straight into that rank's buffer
  // (peer memory) at row gm % route_rows; 0 = local C32
  float* route[16];
  int route_rows;
  long long route_split_stride;   // K-split z of a routed product goes to its own slab, z * stride elements further
};

// TF32 = false: bf16 operands (64 per 128-byte k-block row, UMMA K = 16).  TF32 = true: fp32 operands read as
// TF32 (32 per k-block row, UMMA K = 8) -- the byte geometry of tiles, swizzle and descriptors is identical, so
// the same pipeline serves both; the small GEMMs between the two gene-sized layers use it for operand precision.
template <int BN, bool A_MN, bool B_MN, bool TF32>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using S = GemmSmem<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzled tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full = empty_bar + S::kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent: work units (n tile fastest, then m tile, then K split) are dealt round-robin to the CTAs;
  // two TMEM accumulators let the epilogue of unit i overlap the main loop of unit i+1
  const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + BM - 1) / BM;
  const int num_units = tiles_n * tiles_m * p.splits;
  constexpr int BKE = TF32 ? 32 : 64;   // elements per k-block row (128 bytes)
  constexpr int KI = TF32 ? 8 : 16;     // K of one UMMA
  constexpr int MNB = BKE;              // elements of one 128-byte MN-major group
  const int total_kb = (p.K + BKE - 1) / BKE;
  const int kb_per = (total_kb + p.splits - 1) / p.splits;
  auto unit_coords = [&](int u, int& m0, int& n0, int& z, int& kb0, int& num_kb) {
    n0 = (u % tiles_n) * BN;
    m0 = ((u / tiles_n) % tiles_m) * BM;
    z = u / (tiles_n * tiles_m);
    kb0 = z * kb_per;
    num_kb = max(0, min(total_kb, kb0 + kb_per) - kb0);
  };

  // warp roles: the MMA issuer is the LAST warp (the scheduler favours higher warp ids, and the one issuing
  // thread must never wait behind polling warps), TMA producer next, epilogue warps 0..3
  constexpr int kTmaWarp = 4, kMmaWarp = 5;
  if (warp == kTmaWarp && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // the prologue above overlapped the previous kernel; its results are visible from here on

  if (warp == kTmaWarp) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        int m0, n0, z, kb0, num_kb;
        unit_coords(u, m0, n0, z, kb0, num_kb);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * S::kStageBytes;
          uint8_t* sB = sA + S::kABytes;
          mbar_expect_tx(&full_bar[stage], S::kStageBytes);
          const int k0 = (kb0 + kb) * BKE;
          if (!A_MN) {
            tma_load_2d(sA, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / MNB; ++j) tma_load_2d(sA + j * (BKE * 128), &tmA, &full_bar[stage], m0 + MNB * j, k0);
          }
          if (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / MNB; ++j) tma_load_2d(sB + j * (BKE * 128), &tmB, &full_bar[stage], n0 + MNB * j, k0);
          }
          if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = TF32 ? make_idesc_tf32(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0)
                                      : make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
        int m0, n0, z, kb0, num_kb;
        unit_coords(u, m0, n0, z, kb0, num_kb);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sB = sA + S::kABytes;
#pragma unroll
          for (int k = 0; k < BKE / KI; ++k) {
            // K-major: 8-row groups 1024 B apart, +32 B per K step (16 bf16 / 8 tf32) inside the swizzle span.
            // MN-major: 128-byte MN groups BKE*128 B apart (LBO), 8 K-rows = 1024 B (SBO), +KI*128 B per K step.
            // (MN-major TF32: 4-row swizzle atoms of 512 B, see make_desc_sw128_base32)
            const uint64_t da = A_MN ? (TF32 ? make_desc_sw128_base32(sA + k * (KI * 128), BKE * 128, 512)
                                             : make_desc_sw128(sA + k * (KI * 128), BKE * 128, 1024))
                                     : make_desc_sw128(sA + k * 32, 16, 1024);
            const uint64_t db = B_MN ? (TF32 ? make_desc_sw128_base32(sB + k * (KI * 128), BKE * 128, 512)
                                             : make_desc_sw128(sB + k * (KI * 128), BKE * 128, 1024))
                                     : make_desc_sw128(sB + k * 32, 16, 1024);
            if (TF32) umma_tf32(tmem_d, da, db, idesc, (kb | k) ? 1u : 0u);
            else umma_bf16(tmem_d, da, db, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);   // with num_kb == 0 this arrives immediately (nothing outstanding)
      }
      pdl_trigger();   // every MMA of this CTA is issued: the next kernel in the stream may move in
    }
  } else {
    // ===== epilogue (warps 0..3; TMEM lane group = warp % 4) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float ssq = 0.f;
    int it = 0;
    int chunk_no = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
      int m0, n0, z, kb0, num_kb;
      unit_coords(u, m0, n0, z, kb0, num_kb);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int gm = m0 + row;
      mbar_wait_relaxed(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      // (a routed K-split without k-blocks still has to deliver its slab: zeros)
      for (int c = 0; c < BN / 32 && (num_kb > 0 || p.route_rows > 0); ++c) {
        uint32_t r[32];
        if (num_kb > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        const int gn0 = n0 + c * 32;
        if (p.tma_out || p.route_rows > 0) {
          // f32 chunk [128 rows x 32 cols] -> 128B-swizzled smem (double buffered) -> TMA bulk store (rows and
          // columns outside C are clipped by the tensor map), or -- routed output -- cooperative row stores into the
          // owners' buffers: a warp writes 4 rows x 128 contiguous bytes per instruction, so what crosses NVLink
          // are full 128-byte lines (per-thread 16-byte stores to 32 different rows cost 10x the link time)
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (gn0 + j < p.N) v[j] += __ldg(p.bias + gn0 + j);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.sumsq && gm < p.M) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (gn0 + j < p.N) ssq = fmaf(v[j], v[j], ssq);
          }
          uint8_t* obuf = smem + S::kOutOff + (chunk_no & 1) * (BM * 128);
          if (p.tma_out && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          named_bar_sync(1, 128);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(obuf + row * 128 + ((k ^ (row & 7)) << 4)) =
                make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          if (p.tma_out) fence_proxy_async();
          named_bar_sync(1, 128);
          if (p.tma_out) {
            if (threadIdx.x == 0) {
              tma_store_2d(&tmC, obuf, gn0, m0);
              tma_store_commit();
            }
          } else {
#pragma unroll
            for (int it2 = 0; it2 < 8; ++it2) {
              const int idx = it2 * 128 + (int)threadIdx.x;
              const int rr = idx >> 3, ch = idx & 7;
              const int grow = m0 + rr, gcol = gn0 + ch * 4;
              if (grow < p.M && gcol < p.N) {
                const float4 t = *reinterpret_cast<const float4*>(obuf + rr * 128 + ((ch ^ (rr & 7)) << 4));
                float* dst = p.route[grow / p.route_rows] + (size_t)z * p.route_split_stride +
                             (size_t)(grow % p.route_rows) * p.ldc + gcol;
                if (gcol + 4 <= p.N) {
                  *reinterpret_cast<float4*>(dst) = t;
                } else {
                  const float tv[4] = {t.x, t.y, t.z, t.w};
                  for (int j = 0; j < 4 && gcol + j < p.N; ++j) dst[j] = tv[j];
                }
              }
            }
          }
          ++chunk_no;
          continue;
        }
        if (gm < p.M && gn0 < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          const bool full = (gn0 + 32 <= p.N) && p.vec_ok;
          if (p.splits > 1 && p.route_rows == 0) {
            // split-K: this unit holds a partial product; split 0 also contributes the bias
            float* crow = p.C32 + (size_t)gm * p.ldc + gn0;
            if (p.bias && z == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (gn0 + j < p.N) v[j] += __ldg(p.bias + gn0 + j);
            }
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                atomicAdd(reinterpret_cast<float4*>(crow + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (gn0 + j < p.N) atomicAdd(crow + j, v[j]);
            }
            continue;
          }
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (gn0 + j < p.N) v[j] += __ldg(p.bias + gn0 + j);
          }
          const size_t o = (size_t)gm * p.ldc + gn0;
          if (full) {
            if (p.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(p.C32 + o + j);
                v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.sumsq) {
#pragma unroll
              for (int j = 0; j < 32; ++j) ssq = fmaf(v[j], v[j], ssq);
            }
            if (p.C32) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(p.C32 + o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            if (p.C16) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 t;
                t.x = pack_bf16(v[j], v[j + 1]); t.y = pack_bf16(v[j + 2], v[j + 3]);
                t.z = pack_bf16(v[j + 4], v[j + 5]); t.w = pack_bf16(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(p.C16 + o + j) = t;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (gn0 + j < p.N) {
                float x = v[j];
                if (p.accumulate) x += p.C32[o + j];
                if (p.relu) x = fmaxf(x, 0.f);
                if (p.sumsq) ssq = fmaf(x, x, ssq);
                if (p.C32) p.C32[o + j] = x;
                if (p.C16) p.C16[o + j] = __float2bfloat16(x);
              }
            }
          }
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (p.tma_out && threadIdx.x == 0) tma_store_wait_all();   // smem must outlive the last bulk store
    if (p.sumsq) {
      const double tot = warp_sum((double)ssq);
      if (lane == 0 && tot != 0.0) atomicAdd(p.sumsq, tot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<2 * BN>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int make_tmap(CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* base, uint64_t inner,
                    uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                    CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B);

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows of pitch `ld` elements
int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                   uint32_t box_outer) {
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, inner, outer, ld, box_inner, box_outer);
}
int make_tmap_f32(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                  uint32_t box_outer) {
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, inner, outer, ld, box_inner, box_outer);
}
// f32 in global memory -> TF32 operand tiles.  Data type TFLOAT32: the TMA unit ROUNDS to TF32 on the way into shared
// memory (with plain FLOAT32 the tensor core truncates the low 13 mantissa bits, a systematic -2^-12 bias per
// operand that shrinks every layer's output and shows up as a 6e-4 shift of the loss).
static int make_tmap_tf32(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                          uint32_t box_inner, uint32_t box_outer) {
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, base, inner, outer, ld, box_inner, box_outer);
}
// ... for an MN-major TF32 operand: 128-byte swizzle with 32-byte atoms
static int make_tmap_tf32_mn(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                             uint32_t box_inner, uint32_t box_outer) {
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, base, inner, outer, ld, box_inner, box_outer,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

static int make_tmap(CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* base, uint64_t inner,
                    uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return -3;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (uint64_t)esize};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r, base,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
    return -3;
  }
  return 0;
}

template <int BN, bool A_MN, bool B_MN, bool TF32>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                       cudaStream_t st) {
  using S = GemmSmem<BN>;
  auto kern = gemm_bf16_tc_kernel<BN, A_MN, B_MN, TF32>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) {
      set_error("gemm_bf16_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  const long long units = (long long)((p.N + BN - 1) / BN) * ((p.M + BM - 1) / BM) * p.splits;
  const int grid = (int)(units < sm_budget() ? units : sm_budget());
  launch_pdl(kern, dim3(grid), dim3(kGemmThreads), S::kTotal, st, tmA, tmB, tmC, p);
  return check_launch("gemm_bf16_tc");
}

template <int BN, bool TF32 = false>
static int dispatch_major(int transA, int transB, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const CUtensorMap& tmC, const GemmParams& p, cudaStream_t st) {
  if (!transA && !transB) return launch_gemm<BN, false, false, TF32>(tmA, tmB, tmC, p, st);
  if (!transA && transB) return launch_gemm<BN, false, true, TF32>(tmA, tmB, tmC, p, st);
  if (transA && !transB) return launch_gemm<BN, true, false, TF32>(tmA, tmB, tmC, p, st);
  return launch_gemm<BN, true, true, TF32>(tmA, tmB, tmC, p, st);
}

}  // namespace cmmvae

using namespace cmmvae;

static int gemm_bf16_tc(const void* A, int lda, int transA, const void* Bm, int ldb, int transB, int M, int N, int K,
                        const float* bias, int relu, int accumulate, float* C_f32, void* C_bf16, int ldc,
                        double* sumsq_out, float* const* route, int n_route, int route_rows, void* stream,
                        bool tf32 = false, int n_split = 1, long long split_stride = 0) {
  CMMVAE_REQUIRE(M > 0 && N > 0 && K > 0 && ldc >= N, "gemm_bf16_tc: bad shape M=%d N=%d K=%d ldc=%d", M, N, K, ldc);
  CMMVAE_REQUIRE(C_f32 || C_bf16, "gemm_bf16_tc: no output");
  CMMVAE_REQUIRE(!accumulate || C_f32, "gemm_bf16_tc: accumulate needs C_f32");
  CMMVAE_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)Bm & 15) == 0, "gemm_bf16_tc: operands must be 16-byte aligned");
  const int ea = tf32 ? 4 : 8;   // elements per 16 bytes
  CMMVAE_REQUIRE(lda % ea == 0 && ldb % ea == 0, "gemm_tc: lda/ldb must be multiples of %d (got %d, %d)", ea, lda, ldb);
  const long long tiles256 = (long long)((M + 127) / 128) * ((N + 255) / 256);
  const int BKE = tf32 ? 32 : 64;
  const int total_kb = (K + BKE - 1) / BKE;
  // long-K, few-tile products (dh = dlogits Wout: K = genes) are split along K to fill the 148 SMs
  int splits = 1;
  const int sms = sm_budget();
  // (narrow outputs, N <= 128 -- the output discriminator's first layer -- walk 128-wide tiles and split as well)
  const long long tiles_sel = N > 128 ? tiles256 : (long long)((M + 127) / 128);
  if (!relu && !C_bf16 && !accumulate && C_f32 && total_kb >= 64 && tiles_sel * 2 <= sms && !route) {
    splits = (int)(sms / tiles_sel);
    if (splits > total_kb / 16) splits = total_kb / 16;
    if (splits < 1) splits = 1;
  }
  const int BN = (N > 128 && (tiles256 >= sms || splits > 1)) ? 256 : 128;
  CUtensorMap tmA, tmB;
  int rc;
  // K-major: inner = K, rows = M (or N).  MN-major: inner = M (or N), rows = K, box (128 bytes) x (one k-block).
  auto tmap = tf32 ? make_tmap_tf32 : make_tmap_bf16;
  auto tmap_mn = tf32 ? make_tmap_tf32_mn : make_tmap_bf16;
  if (!transA) rc = tmap(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BKE, BM);
  else rc = tmap_mn(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BKE, BKE);
  if (rc) return rc;
  if (!transB) rc = tmap(&tmB, Bm, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BKE, BN);
  else rc = tmap_mn(&tmB, Bm, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BKE, BKE);
  if (rc) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.relu = relu; p.accumulate = accumulate;
  p.C32 = C_f32; p.C16 = (__nv_bfloat16*)C_bf16; p.ldc = ldc;
  p.vec_ok = (ldc % 8 == 0) && (!C_f32 || ((uintptr_t)C_f32 & 15) == 0) && (!C_bf16 || ((uintptr_t)C_bf16 & 15) == 0);
  p.splits = splits;
  p.sumsq = sumsq_out;
  p.route_rows = 0;
  p.route_split_stride = 0;
  for (int i = 0; i < 16; ++i) p.route[i] = nullptr;
  if (route) {
    CMMVAE_REQUIRE(n_route >= 1 && n_route <= 16 && route_rows > 0 && (long long)n_route * route_rows >= M,
                   "gemm_bf16_tc_routed: %d routes of %d rows do not cover %d rows", n_route, route_rows, M);
    CMMVAE_REQUIRE(n_split >= 1 && (n_split == 1 || split_stride >= (long long)route_rows * ldc),
                   "gemm_bf16_tc_routed: bad split slabs");
    p.route_split_stride = split_stride;
    p.splits = n_split < total_kb ? n_split : total_kb;   // K-split z -> slab z on the owner (no atomics)
    CMMVAE_REQUIRE(!relu && !C_bf16 && !accumulate && !sumsq_out && ldc % 4 == 0,
                   "gemm_bf16_tc_routed: plain f32 output only, ldc a multiple of 4");
    for (int i = 0; i < n_route; ++i) {
      CMMVAE_REQUIRE(route[i] && ((uintptr_t)route[i] & 15) == 0, "gemm_bf16_tc_routed: bad route %d", i);
      p.route[i] = route[i];
    }
    p.route_rows = route_rows;
  }
  CMMVAE_REQUIRE(!sumsq_out || splits == 1, "gemm_bf16_tc: sumsq_out is not available on the split-K path");
  cudaStream_t st = (cudaStream_t)stream;
  if (splits > 1) {
    cudaError_t e = cudaMemset2DAsync(C_f32, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st);
    if (e != cudaSuccess) {
      set_error("gemm_bf16_tc: split-K memset: %s", cudaGetErrorString(e));
      return -2;
    }
  }
  // big f32-only outputs (weight gradients) leave through TMA bulk stores
  CUtensorMap tmC = tmA;
  p.tma_out = (C_f32 && !C_bf16 && !accumulate && splits == 1 && ldc % 4 == 0 && ((uintptr_t)C_f32 & 15) == 0 &&
               (long long)M * N >= (1 << 20) && !route) ? 1 : 0;
  if (p.tma_out)
    if (int rc2 = make_tmap_f32(&tmC, C_f32, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32, BM)) return rc2;
  if (tf32) {
    if (BN == 256) return dispatch_major<256, true>(transA, transB, tmA, tmB, tmC, p, st);
    return dispatch_major<128, true>(transA, transB, tmA, tmB, tmC, p, st);
  }
  if (BN == 256) return dispatch_major<256>(transA, transB, tmA, tmB, tmC, p, st);
  return dispatch_major<128>(transA, transB, tmA, tmB, tmC, p, st);
}

extern "C" int cmmvae_gemm_tf32_tc(const float* A, int lda, int transA, const float* Bm, int ldb, int transB, int M,
                                   int N, int K, const float* bias, int relu, int accumulate, float* C_f32,
                                   void* C_bf16, int ldc, double* sumsq_out, void* stream) {
  return gemm_bf16_tc(A, lda, transA, Bm, ldb, transB, M, N, K, bias, relu, accumulate, C_f32, C_bf16, ldc, sumsq_out,
                      nullptr, 0, 0, stream, true);
}

extern "C" int cmmvae_gemm_bf16_tc(const void* A, int lda, int transA, const void* Bm, int ldb, int transB, int M,
                                   int N, int K, const float* bias, int relu, int accumulate, float* C_f32,
                                   void* C_bf16, int ldc, double* sumsq_out, void* stream) {
  return gemm_bf16_tc(A, lda, transA, Bm, ldb, transB, M, N, K, bias, relu, accumulate, C_f32, C_bf16, ldc, sumsq_out,
                      nullptr, 0, 0, stream);
}

extern "C" int cmmvae_gemm_bf16_tc_routed(const void* A, int lda, int transA, const void* Bm, int ldb, int transB,
                                          int M, int N, int K, int ldc, float* const* route, int n_route,
                                          int route_rows, int n_split, long long split_stride, void* stream) {
  CMMVAE_REQUIRE(route, "gemm_bf16_tc_routed: no routes");
  return gemm_bf16_tc(A, lda, transA, Bm, ldb, transB, M, N, K, nullptr, 0, 0, route[0], nullptr, ldc, nullptr, route,
                      n_route, route_rows, stream, false, n_split, split_stride);
}
