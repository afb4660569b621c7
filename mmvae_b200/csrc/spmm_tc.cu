// K1 / K1b on the tensor pipe: the CSR batch is densified tile by tile INSIDE shared memory (never in
// HBM) straight into the UMMA operand layout, and multiplied with tcgen05.
//
//   forward   Y[B,H]   = X[B,G]   * Wt[G,H]  (+bias)     A = X tile   (K-major,  K = genes)
//   backward  dWt[G,H] = X^T[G,B] * dY[B,H]              A = X^T tile (MN-major, K = cells)
//
// Why: the per-nonzero gather kernel (sparse.cu) moves 2 KB of weight row per non-zero through L1/L2
// (6.3 GB per step at B=1024, d=5%) and is L1/L2-bandwidth bound at ~6% of the HBM roofline of the
// algorithmic bytes.  Above ~1.5% density the dense tile product on the tensor pipe is faster even though
// it multiplies zeros; HBM traffic is the algorithmic minimum: CSR once (per N tile),
// the bf16 weight once, Y once.  Values are rounded to bf16 when staged (the bf16 precision policy).
//
// Persistent CTAs (one per SM) walk 128 x 256 output tiles; 704 threads:
//   warp 20       TMA: B-operand tiles (Wt or dY, MN-major, 4 boxes of 64x64 per stage)
//   warp 21       tcgen05.mma issue (highest warp id = issue priority), two 256-column TMEM accumulators
//   warps 0..15   four producer groups of 4 warps; group g builds the A tile of flat k-blocks f = g (mod 4),
//                 i.e. owns ring stage g: zero the 16 KB tile, scatter the CSR entries of the tile (located
//                 through the per-(cell, 64-gene window) pointer table, prefetched two tiles ahead) into the
//                 128B-swizzled layout, fence.proxy.async, arrive on the stage's full barrier
//   warps 16..19  epilogue: drain accumulator i while tile i+1 is being multiplied
#include "tc.cuh"

namespace cmmvae {

using namespace tc;

constexpr int SBM = 128, SBN = 256, SBK = 64, SSTAGES = 4;
#ifndef SP_E
#define SP_E 8
#endif
constexpr int SNG = 4;             // producer groups (4 warps each); group g builds flat k-blocks f = g (mod SNG)
constexpr int kSpThreads = 64 + 128 * SNG + 128;   // TMA + MMA + producer warps + 4 epilogue warps
constexpr int kSpABytes = SBM * SBK * 2;   // 16 KB
constexpr int kSpBBytes = SBN * SBK * 2;   // 32 KB
constexpr int kSpStage = kSpABytes + kSpBBytes;
constexpr int kSpOutOff = SSTAGES * kSpStage;              // 2 x [128 rows][128 B] f32 chunks, SW128 (backward)
constexpr int kSpBarOff = kSpOutOff + 2 * SBM * 128;
constexpr int kSpSmem = kSpBarOff + 256;

struct SpParams {
  int B, G, H;
  const uint32_t* packed;  // per non-zero: column (low 16 bits) | bf16 value (high 16 bits), CSR order
  const int32_t* tp;       // [ntp][B] (window-major) first CSR position of row b with column >= 64*w
  int ntp;
  const float* bias;   // forward only
  float* out;          // forward: Y[B,H]; backward: dWt[G,H]
  int splits;          // forward split-K over genes
  double* sumsq;       // backward only, optional: += sum of squares of dWt (fused gradient norm)
  int win0, row0;      // backward only: window index of tp's first row / gene index of out's first row (shard views)
  int m_begin, m_end;  // backward only: gene range [m_begin, m_end) computed by this launch (m_begin % 128 == 0)
  // forward only, data parallel: output row gm belongs to rank gm / route_rows and is stored straight into that
  // rank's buffer (peer memory over NVLink) at row gm % route_rows -- the "scatter" of a reduce-scatter done by
  // the epilogue while the next tile is being multiplied.  route_rows == 0: plain local output `out`.
  float* route[16];
  int route_rows;
  long long route_split_stride;   // K-split z of a routed product goes to its own slab, z * stride elements further
};

// tp[w][b] = first position p in row b with col[p] >= 64*w, w = 0..NW (NW = ceil(G/64)); window-major so
// that the producers' reads (consecutive lanes = consecutive cells) coalesce.
// One thread per non-zero: entry i of row b opens every window in (window(i-1), window(i)] (all windows up
// to window(i) when it is the row's first entry); the row's last entry also closes the windows after it.
// Every table cell is written exactly once; rows without entries are filled by tile_ptr64_empty_rows.
// (rbeg / rend: first / one-past-last position of every row; plain CSR passes crow and crow + 1)
__global__ void __launch_bounds__(256) tile_ptr64_kernel(const int32_t* __restrict__ rbeg,
                                                         const int32_t* __restrict__ rend,
                                                         const int32_t* __restrict__ col, int B, int ntp,
                                                         int32_t* __restrict__ tp) {
  pdl_sync();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {      // one CTA per row
    const int s = rbeg[b], e = rend[b];
    if (s == e) {
      for (int w = threadIdx.x; w < ntp; w += blockDim.x) tp[(size_t)w * B + b] = s;
      continue;
    }
    for (int i = s + threadIdx.x; i < e; i += blockDim.x) {
      const int wi = __ldg(col + i) >> 6;
      const int wprev = (i == s) ? -1 : (__ldg(col + i - 1) >> 6);
      for (int w = wprev + 1; w <= wi; ++w) tp[(size_t)w * B + b] = i;
      if (i == e - 1)
        for (int w = wi + 1; w < ntp; ++w) tp[(size_t)w * B + b] = e;
    }
  }
}

// packed[i] = col[i] | bf16(val[i]) << 16   (G <= 65536); one 4-byte record per non-zero
__global__ void csr_pack_kernel(const int32_t* __restrict__ col, const float* __restrict__ val, long long nnz,
                                long long padded, uint32_t* __restrict__ packed,
                                const int32_t* __restrict__ nnz_dev) {
  pdl_sync();
  if (nnz_dev) nnz = min((long long)*nnz_dev, nnz);   // real count from crow[B] (graph replay); `nnz` = capacity
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < padded;
       i += (long long)gridDim.x * blockDim.x) {
    uint32_t r = 0;
    if (i < nnz) {
      const __nv_bfloat16 h = __float2bfloat16(val[i]);
      r = (uint32_t)(col[i] & 0xFFFF) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&h)) << 16);
    }
    packed[i] = r;
  }
}

template <bool BWD>
__global__ void __launch_bounds__(kSpThreads, 1)
spmm_tc_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const SpParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzled tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kSpBarOff);
  uint64_t* empty_bar = full_bar + SSTAGES;
  uint64_t* tmem_full = empty_bar + SSTAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent: units (n tile fastest, then m tile, then K split) are dealt round-robin to the CTAs
  const int Kdim = BWD ? p.B : p.G;
  const int Mdim = BWD ? p.m_end : p.B;
  const int m_base = BWD ? p.m_begin : 0;
  const int tiles_n = (p.H + SBN - 1) / SBN, tiles_m = (Mdim - m_base + SBM - 1) / SBM;
  const int num_units = tiles_n * tiles_m * p.splits;
  const int total_kb = (Kdim + SBK - 1) / SBK;
  const int kb_per = (total_kb + p.splits - 1) / p.splits;
  auto unit_coords = [&](int u, int& m0, int& n0, int& z, int& kb0, int& num_kb) {
    n0 = (u % tiles_n) * SBN;
    m0 = m_base + ((u / tiles_n) % tiles_m) * SBM;
    z = u / (tiles_n * tiles_m);
    kb0 = z * kb_per;
    num_kb = max(0, min(total_kb, kb0 + kb_per) - kb0);
  };

  constexpr int kEpiWarp0 = 4 * SNG, kTmaWarp = 4 * SNG + 4, kMmaWarp = 4 * SNG + 5;   // MMA = highest id
  if (warp == kTmaWarp && lane == 0) {
    prefetch_tmap(&tmB);
    for (int s = 0; s < SSTAGES; ++s) {
      mbar_init(&full_bar[s], 5);   // 1 TMA arrive.expect_tx + 4 producer warps
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<2 * SBN>(tmem_slot);
  // the A tiles start out all zero; the producers keep them that way between occupants (see below)
  for (int i = threadIdx.x; i < SSTAGES * (kSpABytes / 16); i += kSpThreads)
    st_shared_zero16(smem_u32(smem) + (i / (kSpABytes / 16)) * kSpStage + (i % (kSpABytes / 16)) * 16);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // prologue overlapped the previous kernel; its results are visible from here on

  if (warp == kTmaWarp) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        int m0, n0, z, kb0, num_kb;
        unit_coords(u, m0, n0, z, kb0, num_kb);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sB = smem + stage * kSpStage + kSpABytes;
          mbar_expect_tx(&full_bar[stage], kSpBBytes);
#pragma unroll
          for (int j = 0; j < SBN / 64; ++j)
            tma_load_2d(sB + j * (SBK * 128), &tmB, &full_bar[stage], n0 + 64 * j, (kb0 + kb) * SBK);
          if (++stage == SSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(SBM, SBN, BWD ? 1 : 0, 1);
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
        const uint32_t tmem_d = tmem_base + acc * SBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * kSpStage);
          const uint32_t sB = sA + kSpABytes;
#pragma unroll
          for (int k = 0; k < SBK / 16; ++k) {
            const uint64_t da = BWD ? make_desc_sw128(sA + k * 2048, SBK * 128, 1024)
                                    : make_desc_sw128(sA + k * 32, 16, 1024);
            const uint64_t db = make_desc_sw128(sB + k * 2048, SBK * 128, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == SSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
      }
      pdl_trigger();   // every MMA of this CTA is issued: the next kernel in the stream may move in
    }
  } else if (warp < kEpiWarp0) {
    // ===== A-tile producers: SNG groups of 4 warps; group g builds flat k-blocks f = g (mod SNG) =====
    const int group = warp >> 2;                       // 0 .. SNG-1
    const int t = (warp - group * 4) * 32 + lane;      // 0..127 inside the group
    // this thread's 128-byte line of the tile and its swizzle phase
    int line_off, swz;
    if (!BWD) {                       // thread = cell row of the tile; window advances with kb
      line_off = (t >> 3) * 1024 + (t & 7) * 128;
      swz = t & 7;
    } else {                          // thread = (cell kk inside the k-block, gene half h)
      const int kk = t & 63, h = t >> 6;
      line_off = h * (SBK * 128) + (kk >> 3) * 1024 + (kk & 7) * 128;
      swz = kk & 7;
    }
    // cursor over this CTA's flattened (unit, k-block) sequence
    struct Seq { int u, kb, nkb, m0, kb0; };
    auto seq_load = [&](Seq& s) {
      if (s.u < num_units) {
        int n0, z;
        unit_coords(s.u, s.m0, n0, z, s.kb0, s.nkb);
      } else {
        s.nkb = 0; s.m0 = 0; s.kb0 = 0;
      }
    };
    auto seq_advance = [&](Seq& s, int steps) {
      while (steps > 0 && s.u < num_units) {
        const int room = s.nkb - s.kb;
        if (steps < room) { s.kb += steps; return; }
        steps -= room;
        s.u += gridDim.x; s.kb = 0;
        seq_load(s);
      }
    };
    struct Ptr { int q0, q1, win; };
    auto window_ptrs = [&](const Seq& s, Ptr& o) {
      // CSR range [q0,q1) of the entries this thread stages for the k-block at cursor s (coalesced reads)
      o.q0 = o.q1 = 0;
      o.win = 0;
      if (s.u >= num_units) return;
      const int kbg = s.kb0 + s.kb;
      int b, w;
      if (!BWD) { b = s.m0 + t; w = kbg; } else { b = kbg * 64 + (t & 63); w = (s.m0 >> 6) + (t >> 6); }
      if (b < p.B && w < p.ntp - 1) {
        const int wl = BWD ? w - p.win0 : w;
        o.q0 = __ldg(p.tp + (size_t)wl * p.B + b);
        o.q1 = __ldg(p.tp + (size_t)(wl + 1) * p.B + b);
      }
      o.win = w * 64;
    };
    constexpr int E = SP_E;   // packed records prefetched into registers per (thread, window)
    auto load_entries = [&](const Ptr& q, uint32_t (&rec)[E]) {
      const int n = q.q1 - q.q0;
      const uint32_t* src = p.packed + q.q0;
#pragma unroll
      for (int u = 0; u < E; ++u) rec[u] = (u < n) ? __ldg(src + u) : 0u;
    };
    // Software pipeline per group over its flat indices f = group, group+2, ...: pointers for f+4 and
    // entries for f+2 are requested while tile f is built, so every global load has a full iteration to
    // land.  Three pointer slots and two entry slots rotate by NAME (six body instances per trip) so no
    // register moves touch in-flight loads.
    int total_f = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      int m0, n0, z, kb0, nkb;
      unit_coords(u, m0, n0, z, kb0, nkb);
      total_f += nkb;
    }
    Seq cur_s{(int)blockIdx.x, 0, 0, 0, 0};
    seq_load(cur_s);
    seq_advance(cur_s, group);            // cursor at flat index f = group
    Seq fut_s = cur_s;                    // will run 4 flat steps (two group iterations) ahead
    Ptr P0, P1, P2;
    uint32_t E0[E], E1[E];
    window_ptrs(fut_s, P0);
    seq_advance(fut_s, SNG);
    window_ptrs(fut_s, P1);
    load_entries(P0, E0);
    static_assert(SNG == SSTAGES, "group g must own ring stage g (its threads keep the stage's un-scatter state)");
    constexpr int EW = (SP_E + 3) / 4;   // entry byte offsets, four per register
    uint32_t prev_off[EW];
#pragma unroll
    for (int i = 0; i < EW; ++i) prev_off[i] = 0;
    int prev_n = 0, prev_a0 = 0, prev_aw = 0;
    const uint32_t smem_base = smem_u32(smem);
    auto body = [&](int f, const Ptr& cur, const Ptr& nxt, Ptr& fut, const uint32_t (&ec)[E], uint32_t (&en)[E]) {
      if (f >= total_f) return;
      const int stage = f % SSTAGES;
      const uint32_t phase = (f / SSTAGES) & 1;
      seq_advance(fut_s, SNG);
      window_ptrs(fut_s, fut);
      load_entries(nxt, en);
      mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
      const uint32_t line = smem_base + stage * kSpStage + line_off;
      // The tile is all zeros except the entries this group scattered for the stage's previous occupant, and a
      // thread only ever writes its own 128-byte line: take the previous entries back out (byte offsets kept
      // packed in two registers), then put the new ones in.  No 16 KB re-zeroing, no barrier inside the group.
#pragma unroll
      for (int u = 0; u < E; ++u)
        st_shared_u16_if(line + ((prev_off[u >> 2] >> (8 * (u & 3))) & 0xFFu), 0u, u < prev_n);
      for (int q = prev_a0 + E; q < prev_a0 + prev_n; ++q) {   // previous window had more than E entries
        const int c = (int)(__ldg(p.packed + q) & 0xFFFFu) - prev_aw;
        st_shared_u16_if(line + ((((c >> 3) ^ swz) << 4) | ((c & 7) << 1)), 0u, true);
      }
      const int a0 = cur.q0, aw = cur.win;
      const int n = cur.q1 - cur.q0;
      uint32_t offs[EW];
#pragma unroll
      for (int i = 0; i < EW; ++i) offs[i] = 0;
#pragma unroll
      for (int u = 0; u < E; ++u) {
        const int c = (int)(ec[u] & 0xFFFFu) - aw;
        const uint32_t off = (uint32_t)((((c >> 3) ^ swz) << 4) | ((c & 7) << 1)) & 0x7Fu;
        st_shared_u16_if(line + off, ec[u] >> 16, u < n);
        offs[u >> 2] |= off << (8 * (u & 3));
      }
      for (int q = a0 + E; q < a0 + n; ++q) {   // windows with more than E entries (dense batches)
        const uint32_t r = __ldg(p.packed + q);
        const int c = (int)(r & 0xFFFFu) - aw;
        st_shared_u16_if(line + ((((c >> 3) ^ swz) << 4) | ((c & 7) << 1)), r >> 16, true);
      }
#pragma unroll
      for (int i = 0; i < EW; ++i) prev_off[i] = offs[i];
      prev_n = n; prev_a0 = a0; prev_aw = aw;
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    };
    for (int f = group; f < total_f; f += 6 * SNG) {
      body(f, P0, P1, P2, E0, E1);
      body(f + SNG, P1, P2, P0, E1, E0);
      body(f + 2 * SNG, P2, P0, P1, E0, E1);
      body(f + 3 * SNG, P0, P1, P2, E1, E0);
      body(f + 4 * SNG, P1, P2, P0, E0, E1);
      body(f + 5 * SNG, P2, P0, P1, E1, E0);
    }
  } else {
    // ===== the last 4 warps: epilogue (TMEM lane group = warp % 4): drain unit i while unit i+1 is computed =====
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    float ssq = 0.f;
    int it = 0;
    int chunk_no = 0;
    const int et = (warp - kEpiWarp0) * 32 + lane;   // 0..127 among the epilogue threads
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
      int m0, n0, z, kb0, num_kb;
      unit_coords(u, m0, n0, z, kb0, num_kb);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int gm = m0 + row;
      mbar_wait_relaxed(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      // (a routed K-split without k-blocks -- tiny gene ranges -- still has to deliver its slab: zeros)
      const bool routed = !BWD && p.route_rows > 0;
      for (int c = 0; c < SBN / 32 && (num_kb > 0 || routed); ++c) {
        uint32_t r[32];
        if (num_kb > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * SBN + c * 32), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        const int gn0 = n0 + c * 32;
        if (BWD) {
          // dWt chunk [128 genes x 32 cols] f32 -> 128B-swizzled smem (double buffered) -> TMA bulk store;
          // gene rows beyond G / columns beyond H are clipped by the tensor map
          if (p.sumsq && gm < Mdim) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (gn0 + j < p.H) ssq = fmaf(__uint_as_float(r[j]), __uint_as_float(r[j]), ssq);
          }
          uint8_t* obuf = smem + kSpOutOff + (chunk_no & 1) * (SBM * 128);
          if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          named_bar_sync(15, 128);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<uint4*>(obuf + row * 128 + ((k ^ (row & 7)) << 4)) =
                make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
          fence_proxy_async();
          named_bar_sync(15, 128);
          if (et == 0) {
            tma_store_2d(&tmC, obuf, gn0, m0 - p.row0);
            tma_store_commit();
          }
          ++chunk_no;
          continue;
        }
        if (routed) {
          // partial sums for the cells' owners: stage the [128 cells x 32 cols] chunk in swizzled smem, then the
          // four epilogue warps store it row-wise -- 4 rows x 128 contiguous bytes per warp instruction -- so
          // that full 128-byte lines cross NVLink
          uint8_t* obuf = smem + kSpOutOff + (chunk_no & 1) * (SBM * 128);
          named_bar_sync(15, 128);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<uint4*>(obuf + row * 128 + ((k ^ (row & 7)) << 4)) =
                make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
          named_bar_sync(15, 128);
#pragma unroll
          for (int it2 = 0; it2 < 8; ++it2) {
            const int idx = it2 * 128 + et;
            const int rr = idx >> 3, ch = idx & 7;
            const int grow = m0 + rr, gcol = gn0 + ch * 4;
            if (grow < Mdim && gcol < p.H) {     // H % 8 == 0: a 4-column group is inside or outside as a whole
              const uint4 t = *reinterpret_cast<const uint4*>(obuf + rr * 128 + ((ch ^ (rr & 7)) << 4));
              float* dst = p.route[grow / p.route_rows] + (size_t)z * p.route_split_stride +
                           (size_t)(grow % p.route_rows) * p.H + gcol;
              *reinterpret_cast<uint4*>(dst) = t;
            }
          }
          ++chunk_no;
          continue;
        }
        if (gm < Mdim && gn0 < p.H) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          float* orow = (!BWD && p.route_rows > 0)
                            ? p.route[gm / p.route_rows] + (size_t)z * p.route_split_stride +
                                  (size_t)(gm % p.route_rows) * p.H + gn0
                            : p.out + (size_t)gm * p.H + gn0;
          const bool full = gn0 + 32 <= p.H;
          if (!BWD && p.bias && z == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (gn0 + j < p.H) v[j] += __ldg(p.bias + gn0 + j);
          }
          if (!BWD && p.splits > 1 && p.route_rows == 0) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                atomicAdd(reinterpret_cast<float4*>(orow + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
            } else {
              for (int j = 0; j < 32; ++j)
                if (gn0 + j < p.H) atomicAdd(orow + j, v[j]);
            }
          } else if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(orow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (BWD && p.sumsq) {
#pragma unroll
              for (int j = 0; j < 32; ++j) ssq = fmaf(v[j], v[j], ssq);
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (gn0 + j < p.H) {
                orow[j] = v[j];
                if (BWD && p.sumsq) ssq = fmaf(v[j], v[j], ssq);
              }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (BWD && et == 0) tma_store_wait_all();   // smem must outlive the last bulk store
    if (BWD && p.sumsq) {
      const double tot = warp_sum((double)ssq);
      if (lane == 0 && tot != 0.0) atomicAdd(p.sumsq, tot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<2 * SBN>(tmem_base);
}

int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                   uint32_t box_outer);
int make_tmap_f32(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                  uint32_t box_outer);

template <bool BWD>
static int launch_spmm_tc(const CUtensorMap& tm, const CUtensorMap& tmC, const SpParams& p, dim3 grid,
                          cudaStream_t st) {
  auto kern = spmm_tc_kernel<BWD>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem);
    if (e != cudaSuccess) {
      set_error("spmm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  launch_pdl(kern, dim3(grid), dim3(kSpThreads), kSpSmem, st, tm, tmC, p);
  return check_launch(BWD ? "csr_linear_bwd_w_tc" : "csr_linear_fwd_tc");
}

}  // namespace cmmvae

using namespace cmmvae;

extern "C" size_t cmmvae_csr_tile_ptr_bytes(int B, int G) {
  return sizeof(int32_t) * (size_t)B * (size_t)((G + 63) / 64 + 1);
}
extern "C" size_t cmmvae_csr_packed_bytes(long long nnz) { return sizeof(uint32_t) * (size_t)((nnz + 3) / 4 * 4 + 4); }

static int csr_tile_ptr_rows(const int32_t* rbeg, const int32_t* rend, const int32_t* col, const float* val, int B,
                             int G, long long nnz, int32_t* tile_ptr, void* packed, void* stream,
                             const int32_t* nnz_dev = nullptr);

// nnz is read on the device from crow[B]; `cap` records are written (zeros beyond nnz): the call can be replayed from
// a captured CUDA graph for batches of different density staged at the same addresses
extern "C" int cmmvae_csr_tile_ptr_dyn(const int32_t* crow, const int32_t* col, const float* val, int B, int G,
                                       long long cap, int32_t* tile_ptr, void* packed, void* stream) {
  return csr_tile_ptr_rows(crow, crow + 1, col, val, B, G, cap, tile_ptr, packed, stream, crow + B);
}

extern "C" int cmmvae_csr_tile_ptr(const int32_t* crow, const int32_t* col, const float* val, int B, int G,
                                   long long nnz, int32_t* tile_ptr, void* packed, void* stream) {
  return csr_tile_ptr_rows(crow, crow + 1, col, val, B, G, nnz, tile_ptr, packed, stream);
}

extern "C" int cmmvae_csr_tile_ptr_rows(const int32_t* row_begin, const int32_t* row_end, const int32_t* col,
                                        const float* val, int B, int G, long long n_records, int32_t* tile_ptr,
                                        void* packed, void* stream) {
  return csr_tile_ptr_rows(row_begin, row_end, col, val, B, G, n_records, tile_ptr, packed, stream);
}

static int csr_tile_ptr_rows(const int32_t* crow, const int32_t* crow_end, const int32_t* col, const float* val, int B,
                             int G, long long nnz, int32_t* tile_ptr, void* packed, void* stream,
                             const int32_t* nnz_dev) {
  CMMVAE_REQUIRE(B > 0 && G > 0 && tile_ptr && packed, "csr_tile_ptr: bad arguments");
  CMMVAE_REQUIRE(G <= 65536, "csr_tile_ptr: packed records hold 16-bit gene ids (G=%d)", G);
  CMMVAE_REQUIRE(((uintptr_t)packed & 15) == 0, "csr_tile_ptr: packed must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int ntp = (G + 63) / 64 + 1;
  int blocks = B < 148 * 16 ? B : 148 * 16;   // one CTA per row
  launch_pdl(tile_ptr64_kernel, dim3(blocks), dim3(256), 0, st, crow, crow_end, col, B, ntp, tile_ptr);
  if (int rc = check_launch("csr_tile_ptr")) return rc;
  long long want;
  const long long padded = (nnz + 3) / 4 * 4 + 4;
  want = (padded + 255) / 256;
  blocks = (int)(want < 148 * 16 ? want : 148 * 16);
  launch_pdl(csr_pack_kernel, dim3(blocks), dim3(256), 0, st, col, val, nnz, padded, (uint32_t*)packed, nnz_dev);
  return check_launch("csr_pack");
}

static int spmm_fwd(const void* packed, const int32_t* tile_ptr, int B, int G, int H, const void* Wt_bf16,
                    const float* bias, float* Y, float* const* route, int n_route, int route_rows, int n_split,
                    long long split_stride, void* stream) {
  CMMVAE_REQUIRE(B > 0 && G > 0 && H > 0 && H % 8 == 0, "csr_linear_fwd_tc: bad shape (H must be a multiple of 8)");
  CMMVAE_REQUIRE(((uintptr_t)Wt_bf16 & 15) == 0 && ((uintptr_t)Y & 15) == 0, "csr_linear_fwd_tc: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  SpParams p;
  p.B = B; p.G = G; p.H = H; p.packed = (const uint32_t*)packed; p.tp = tile_ptr; p.ntp = (G + 63) / 64 + 1;
  p.bias = bias; p.out = Y; p.sumsq = nullptr; p.m_begin = 0; p.m_end = B; p.win0 = 0; p.row0 = 0;
  p.route_rows = 0;
  p.route_split_stride = 0;
  for (int i = 0; i < 16; ++i) p.route[i] = nullptr;
  if (route) {
    CMMVAE_REQUIRE(n_route >= 1 && n_route <= 16 && route_rows > 0 && (long long)n_route * route_rows >= B,
                   "csr_linear_fwd_tc_routed: %d routes of %d rows do not cover %d rows", n_route, route_rows, B);
    CMMVAE_REQUIRE(n_split >= 1 && (n_split == 1 || split_stride >= (long long)route_rows * H),
                   "csr_linear_fwd_tc_routed: bad split slabs");
    p.route_split_stride = split_stride;
    for (int i = 0; i < n_route; ++i) {
      CMMVAE_REQUIRE(route[i] && ((uintptr_t)route[i] & 15) == 0, "csr_linear_fwd_tc_routed: bad route %d", i);
      p.route[i] = route[i];
    }
    p.route_rows = route_rows;
  }
  const int tiles = ((B + SBM - 1) / SBM) * ((H + SBN - 1) / SBN);
  const int total_kb = (G + SBK - 1) / SBK;
  const int sms = sm_budget();
  int splits = tiles >= sms ? 1 : sms / tiles;
  if (splits > total_kb / 8) splits = total_kb / 8;
  if (splits < 1) splits = 1;
  if (route) {
    // every partial tile goes to its owner exactly once: K-split z into slab z (the owner sums slabs anyway)
    splits = n_split;
    if (splits > total_kb) splits = total_kb;
  }
  p.splits = splits;
  CUtensorMap tm;
  if (int rc = make_tmap_bf16(&tm, Wt_bf16, (uint64_t)H, (uint64_t)G, (uint64_t)H, 64, SBK)) return rc;
  if (splits > 1 && !route) cudaMemsetAsync(Y, 0, sizeof(float) * (size_t)B * H, st);
  const int units = tiles * splits;
  dim3 grid(units < sms ? units : sms);
  return launch_spmm_tc<false>(tm, tm, p, grid, st);
}

extern "C" int cmmvae_csr_linear_fwd_tc(const void* packed, const int32_t* tile_ptr, int B, int G, int H,
                                        const void* Wt_bf16, const float* bias, float* Y, void* stream) {
  return spmm_fwd(packed, tile_ptr, B, G, H, Wt_bf16, bias, Y, nullptr, 0, 0, 1, 0, stream);
}

extern "C" int cmmvae_csr_linear_fwd_tc_routed(const void* packed, const int32_t* tile_ptr, int B, int G, int H,
                                               const void* Wt_bf16, float* const* route, int n_route,
                                               int route_rows, int n_split, long long split_stride, void* stream) {
  CMMVAE_REQUIRE(route, "csr_linear_fwd_tc_routed: no routes");
  return spmm_fwd(packed, tile_ptr, B, G, H, Wt_bf16, nullptr, route[0], route, n_route, route_rows, n_split,
                  split_stride, stream);
}

static int spmm_bwd_w(const void* packed, const int32_t* tile_ptr, int B, int G, int H, const void* dY_bf16,
                      float* dWt, int g_begin, int g_end, bool shard_view, double* sumsq_out, void* stream) {
  CMMVAE_REQUIRE(B > 0 && G > 0 && H > 0 && H % 8 == 0, "csr_linear_bwd_w_tc: bad shape (H must be a multiple of 8)");
  CMMVAE_REQUIRE(((uintptr_t)dY_bf16 & 15) == 0 && ((uintptr_t)dWt & 15) == 0, "csr_linear_bwd_w_tc: alignment");
  SpParams p;
  p.B = B; p.G = G; p.H = H; p.packed = (const uint32_t*)packed; p.tp = tile_ptr; p.ntp = (G + 63) / 64 + 1;
  if (g_end <= 0 || g_end > G) g_end = G;
  CMMVAE_REQUIRE(g_begin >= 0 && g_begin < g_end && g_begin % SBM == 0 && (g_end == G || g_end % SBM == 0),
                 "csr_linear_bwd_w_tc: gene range [%d,%d) must be 128-aligned", g_begin, g_end);
  p.bias = nullptr; p.out = dWt; p.splits = 1; p.sumsq = sumsq_out; p.m_begin = g_begin; p.m_end = g_end;
  p.route_rows = 0;
  p.route_split_stride = 0;
  for (int i = 0; i < 16; ++i) p.route[i] = nullptr;
  p.win0 = shard_view ? g_begin / 64 : 0;
  p.row0 = shard_view ? g_begin : 0;
  CUtensorMap tm;
  if (int rc = make_tmap_bf16(&tm, dY_bf16, (uint64_t)H, (uint64_t)B, (uint64_t)H, 64, SBK)) return rc;
  const int units = ((H + SBN - 1) / SBN) * ((g_end - g_begin + SBM - 1) / SBM);
  dim3 grid(units < sm_budget() ? units : sm_budget());
  CUtensorMap tmC;
  if (int rc = make_tmap_f32(&tmC, dWt, (uint64_t)H, (uint64_t)(g_end - p.row0), (uint64_t)H, 32, SBM)) return rc;
  return launch_spmm_tc<true>(tm, tmC, p, grid, (cudaStream_t)stream);
}

extern "C" int cmmvae_csr_linear_bwd_w_tc(const void* packed, const int32_t* tile_ptr, int B, int G, int H,
                                          const void* dY_bf16, float* dWt, int g_begin, int g_end,
                                          double* sumsq_out, void* stream) {
  return spmm_bwd_w(packed, tile_ptr, B, G, H, dY_bf16, dWt, g_begin, g_end, false, sumsq_out, stream);
}

extern "C" int cmmvae_csr_linear_bwd_w_tc_shard(const void* packed, const int32_t* tile_ptr_shard, int B, int G,
                                                int H, const void* dY_bf16, float* dWt_shard, int g_begin,
                                                int g_end, double* sumsq_out, void* stream) {
  return spmm_bwd_w(packed, tile_ptr_shard, B, G, H, dY_bf16, dWt_shard, g_begin, g_end, true, sumsq_out, stream);
}
