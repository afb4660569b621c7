// Data-parallel exchange over NVLink / NVSwitch PEER MEMORY (no NCCL on the data path).
//
// Every rank owns a set of symmetric buffers (same layout on every GPU, mapped into all peers through CUDA IPC);
// an exchange is "the producer STORES its piece straight into every consumer's buffer, then raises a flag there".
// Consumers spin on their local flags (system-scope acquire) and read local memory.  The big pieces -- partial
// sums of the gene-sharded first layer and of dh -- are stored by the EPILOGUE of the kernel that computes them
// (spmm_tc.cu / gemm_tc.cu take a PeerRoute), so transfer and math overlap tile by tile and no communication
// kernel ever occupies SMs.  The reduction over ranks happens in the consumer (slab_sum: 8 slabs of 4 MB are a
// few microseconds of HBM traffic).
//
//   cmmvae_peer_push    src -> slot of every peer (16-byte vector stores), then flags        (all-gather piece)
//   cmmvae_peer_signal  flags only (after a kernel whose epilogue already stored to the peers)
//   cmmvae_peer_wait    spin until the local flags of a channel reach `step`
//   cmmvae_slab_sum     out = sum over slabs (+ bias), f32 (+ bf16 copy)                   (the reduce of reduce-scatter)
//   cmmvae_csr_scatter_shards  all-to-all of the batch: piece q (gene shard q) of every row goes straight into rank
//                       q's buffer, columns rebased to the shard -- bit-exact index handling
//   cmmvae_slab_rows    row ranges over the received slabs, in place
#include "common.cuh"

namespace cmmvae {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
  void* p[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

// grid (chunks, peers): block (c, p) copies 16-byte words c, c+chunks, ... of src into peer p's slot; the last
// block to finish (device-scope ticket) raises this rank's flag on every peer
__global__ void __launch_bounds__(256) peer_push_kernel(const uint4* __restrict__ src, long long n16, PeerPtrs dst,
                                                        PeerPtrs flags, int n_peers, uint32_t step,
                                                        unsigned int* __restrict__ ticket,
                                                        const uint32_t* __restrict__ step_dev) {
  pdl_sync();
  if (step_dev) step = *step_dev;    // step number kept in device memory (the launch is replayed from a CUDA graph)
  uint4* out = reinterpret_cast<uint4*>(dst.p[blockIdx.y]);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    out[i] = __ldg(src + i);
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *ticket = 0u;
    __threadfence_system();
    if (threadIdx.x < n_peers && flags.p[threadIdx.x])
      st_release_sys_u32(reinterpret_cast<uint32_t*>(flags.p[threadIdx.x]), step);
  }
}

__global__ void peer_signal_kernel(PeerPtrs flags, int n_peers, uint32_t step, const uint32_t* __restrict__ step_dev) {
  pdl_sync();   // returns once the producer grid has completed and its (peer) stores are visible
  if (step_dev) step = *step_dev;
  __threadfence_system();
  if (threadIdx.x < n_peers) st_release_sys_u32(reinterpret_cast<uint32_t*>(flags.p[threadIdx.x]), step);
}

// NOT a programmatic-launch citizen on purpose: the kernel never triggers its dependents early.  A successor that
// moved onto the SMs while this kernel spins (and then blocked in griddepcontrol.wait) could hold the registers /
// shared memory that a producer kernel of ANOTHER stream of this GPU -- the one a peer is waiting for -- needs:
// a distributed deadlock.  Successors therefore launch only once the flags have arrived.
__global__ void peer_wait_kernel(const uint32_t* __restrict__ local_flags, int n_peers, uint32_t step,
                                 const uint32_t* __restrict__ step_dev) {
  pdl_wait();
  if (step_dev) step = *step_dev;
  if (threadIdx.x < n_peers) {
    // flags only grow; (int) difference keeps the comparison valid across a 2^32 wrap
    while ((int)(ld_acquire_sys_u32(local_flags + threadIdx.x) - step) < 0) __nanosleep(100);
  }
  __syncthreads();
  __threadfence_system();
}

// out[i] = sum_s slabs[s * stride + i] (+ bias[i % H]);  n % 4 == 0, H % 4 == 0
__global__ void __launch_bounds__(256) slab_sum_kernel(const float4* __restrict__ slabs, int n_slabs,
                                                       long long stride4, long long n4,
                                                       const float* __restrict__ bias, int H,
                                                       float4* __restrict__ out32, uint2* __restrict__ out16) {
  pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = slabs[i];
    for (int s = 1; s < n_slabs; ++s) {
      const float4 b = slabs[(long long)s * stride4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + (i * 4) % H));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (out32) out32[i] = a;
    if (out16) out16[i] = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
  }
}

// ---- all-to-all of the batch by gene shard ---------------------------------------------------------------------------
// Instead of gathering every rank's whole batch everywhere (N x the batch per rank), each rank cuts ITS rows into the
// N gene shards and stores piece q straight into rank q's buffer (slab = source rank): crow (B+1), then the entries
// with columns rebased to the shard.  Rows are sorted, so piece q of a row is the contiguous range between two lower
// bounds.  The receiver moves nothing: it only derives row ranges / window pointers over the slabs in place.
//   count: thread per (dest, row)      scan: one CTA per dest (crow written to the peer + a local copy of the offsets)
//   copy:  warp per (dest, row), 128 contiguous bytes per store instruction into peer memory
__global__ void __launch_bounds__(256) scatter_count_kernel(const int32_t* __restrict__ crow,
                                                            const int32_t* __restrict__ col, int B, int n_dst, int per,
                                                            int32_t* __restrict__ cnt, int32_t* __restrict__ start) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n_dst) return;
  const int q = i / B, b = i % B;
  const int r0 = crow[b], r1 = crow[b + 1];
  const int g0 = q * per, g1 = g0 + per;
  int lo = r0, hi = r1;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (__ldg(col + m) < g0) lo = m + 1; else hi = m; }
  const int a = lo;
  hi = r1;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (__ldg(col + m) < g1) lo = m + 1; else hi = m; }
  cnt[i] = lo - a;
  start[i] = a;
}

__global__ void __launch_bounds__(1024) scatter_scan_kernel(const int32_t* __restrict__ cnt, int B, int cap,
                                                            PeerPtrs dst_crow, int32_t* __restrict__ offs,
                                                            int32_t* __restrict__ info) {
  pdl_sync();
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int q = blockIdx.x;
  int32_t* crow_out = reinterpret_cast<int32_t*>(dst_crow.p[q]);
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < B; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < B ? cnt[q * B + i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int excl = carry + (w ? warp_tot[w - 1] : 0) + x - v;
    if (i < B) {
      const int c = min(excl, cap);   // on overflow the tail rows come out empty (flagged), never out of bounds
      crow_out[i] = c;
      offs[q * B + i] = c;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    crow_out[B] = min(carry, cap);
    atomicMax(&info[0], carry);              // densest piece
    if (carry > cap) atomicExch(&info[1], 1);   // overflow
  }
}

__global__ void __launch_bounds__(256) scatter_copy_kernel(const int32_t* __restrict__ col, const float* __restrict__ val,
                                                           int B, int n_dst, int per, int cap,
                                                           const int32_t* __restrict__ cnt,
                                                           const int32_t* __restrict__ start,
                                                           const int32_t* __restrict__ offs, PeerPtrs dst_col,
                                                           PeerPtrs dst_val) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int n_rows = B * n_dst;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_rows; i += (gridDim.x * blockDim.x) >> 5) {
    const int q = i / B;
    int32_t* oc = reinterpret_cast<int32_t*>(dst_col.p[q]);
    float* ov = reinterpret_cast<float*>(dst_val.p[q]);
    const int n = cnt[i], a = start[i], o = offs[i], g0 = q * per;
    for (int k = lane; k < n; k += 32)
      if (o + k < cap) {
        oc[o + k] = __ldg(col + a + k) - g0;
        ov[o + k] = __ldg(val + a + k);
      }
  }
  __threadfence_system();
}

// row ranges of the N received slabs as positions in ONE array that spans all slabs (gaps between slabs are never
// addressed): begin/end[s * B + b] = s * slab_elems + crow_s[b], crow_s[b + 1]
__global__ void __launch_bounds__(256) slab_rows_kernel(const uint8_t* __restrict__ base, long long slab_bytes, int B,
                                                        int n_src, int32_t* __restrict__ rbeg,
                                                        int32_t* __restrict__ rend) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n_src) return;
  const int s = i / B, b = i % B;
  const int32_t* crow = reinterpret_cast<const int32_t*>(base + (long long)s * slab_bytes);
  const int off = (int)((long long)s * (slab_bytes / 4));
  rbeg[i] = off + crow[b];
  rend[i] = off + crow[b + 1];
}

// recon of THIS rank's cells and the squared gradient norm of a gene-sharded group, from the scalar slabs all
// ranks pushed: slab s = [loss_part_s[0..N) | shard_sumsq_s]  (doubles).  out_recon = sum_s slab_s[rank];
// out_norm += sum_s shard_sumsq_s
__global__ void dp_scalars_kernel(const double* __restrict__ slabs, int n_src, int stride, int rank,
                                  double* __restrict__ out_recon, double* __restrict__ out_norm) {
  pdl_sync();
  if (threadIdx.x == 0) {
    double r = 0.0, q = 0.0;
    for (int s = 0; s < n_src; ++s) {
      r += slabs[s * stride + rank];
      q += slabs[s * stride + n_src];
    }
    *out_recon = r;
    *out_norm += q;
  }
}

}  // namespace cmmvae

using namespace cmmvae;

static int fill_ptrs(PeerPtrs& out, void* const* ptrs, int n, const char* what) {
  CMMVAE_REQUIRE(n >= 1 && n <= kMaxPeers && ptrs, "%s: 1..%d peers", what, kMaxPeers);
  for (int i = 0; i < kMaxPeers; ++i) out.p[i] = i < n ? ptrs[i] : nullptr;
  return 0;
}

extern "C" int cmmvae_peer_push(const void* src, long long nbytes, void* const* dst_slots, void* const* peer_flags,
                                int n_peers, unsigned int step, const unsigned int* step_dev, unsigned int* ticket,
                                void* stream) {
  CMMVAE_REQUIRE(nbytes >= 0 && nbytes % 16 == 0 && ((uintptr_t)src & 15) == 0 && ticket,
                 "peer_push: 16-byte aligned source and size");
  PeerPtrs d, f;
  if (int rc = fill_ptrs(d, dst_slots, n_peers, "peer_push")) return rc;
  if (peer_flags) {
    if (int rc = fill_ptrs(f, peer_flags, n_peers, "peer_push")) return rc;
  } else {
    for (int i = 0; i < kMaxPeers; ++i) f.p[i] = nullptr;   // a piece of a multi-part push: flags come with the last part
  }
  for (int i = 0; i < n_peers; ++i) CMMVAE_REQUIRE(((uintptr_t)d.p[i] & 15) == 0, "peer_push: unaligned slot");
  const long long n16 = nbytes / 16;
  long long want = (n16 + 255) / 256;
  const int chunks = (int)(want < 1 ? 1 : (want > 32 ? 32 : want));   // a few CTAs per peer saturate the links
  launch_pdl(peer_push_kernel, dim3(chunks, n_peers), dim3(256), 0, (cudaStream_t)stream, (const uint4*)src, n16, d, f,
             n_peers, (uint32_t)step, ticket, (const uint32_t*)step_dev);
  return check_launch("peer_push");
}

extern "C" int cmmvae_peer_signal(void* const* peer_flags, int n_peers, unsigned int step,
                                  const unsigned int* step_dev, void* stream) {
  PeerPtrs f;
  if (int rc = fill_ptrs(f, peer_flags, n_peers, "peer_signal")) return rc;
  launch_pdl(peer_signal_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, f, n_peers, (uint32_t)step,
             (const uint32_t*)step_dev);
  return check_launch("peer_signal");
}

extern "C" int cmmvae_peer_wait(const void* local_flags, int n_peers, unsigned int step, const unsigned int* step_dev,
                                void* stream) {
  CMMVAE_REQUIRE(local_flags && n_peers >= 1 && n_peers <= kMaxPeers, "peer_wait: bad arguments");
  launch_pdl(peer_wait_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (const uint32_t*)local_flags, n_peers,
             (uint32_t)step, (const uint32_t*)step_dev);
  return check_launch("peer_wait");
}

extern "C" int cmmvae_slab_sum(const float* slabs, int n_slabs, long long slab_stride, long long n, const float* bias,
                               int H, float* out_f32, void* out_bf16, void* stream) {
  CMMVAE_REQUIRE(n_slabs >= 1 && n > 0 && n % 4 == 0 && slab_stride % 4 == 0 && (!bias || H % 4 == 0),
                 "slab_sum: sizes must be multiples of 4");
  CMMVAE_REQUIRE((((uintptr_t)slabs | (uintptr_t)out_f32) & 15) == 0 && ((uintptr_t)out_bf16 & 7) == 0,
                 "slab_sum: alignment");
  const long long n4 = n / 4;
  long long want = (n4 + 255) / 256;
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
  launch_pdl(slab_sum_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const float4*)slabs, n_slabs,
             slab_stride / 4, n4, bias, H, (float4*)out_f32, (uint2*)out_bf16);
  return check_launch("slab_sum");
}

extern "C" int cmmvae_dp_scalars(const double* slabs, int n_src, int stride, int rank, double* out_recon,
                                 double* out_norm, void* stream) {
  CMMVAE_REQUIRE(slabs && n_src >= 1 && stride > n_src && rank >= 0 && rank < n_src, "dp_scalars: bad arguments");
  launch_pdl(dp_scalars_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, slabs, n_src, stride, rank, out_recon,
             out_norm);
  return check_launch("dp_scalars");
}

extern "C" int cmmvae_csr_scatter_shards(const int32_t* crow, const int32_t* col, const float* val, int B, int n_dst,
                                         int per, int cap, void* const* dst_crow, void* const* dst_col,
                                         void* const* dst_val, int32_t* cnt, int32_t* start, int32_t* offs,
                                         int32_t* info, void* stream) {
  CMMVAE_REQUIRE(crow && col && val && B > 0 && per > 0 && cap > 0 && cnt && start && offs && info,
                 "csr_scatter_shards: bad arguments");
  PeerPtrs pc, pl, pv;
  if (int rc = fill_ptrs(pc, dst_crow, n_dst, "csr_scatter_shards")) return rc;
  if (int rc = fill_ptrs(pl, dst_col, n_dst, "csr_scatter_shards")) return rc;
  if (int rc = fill_ptrs(pv, dst_val, n_dst, "csr_scatter_shards")) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = B * n_dst;
  cudaMemsetAsync(info, 0, 2 * sizeof(int32_t), st);
  launch_pdl(scatter_count_kernel, dim3((rows + 255) / 256), dim3(256), 0, st, crow, col, B, n_dst, per, cnt, start);
  if (int rc = check_launch("scatter_count")) return rc;
  launch_pdl(scatter_scan_kernel, dim3(n_dst), dim3(1024), 0, st, (const int32_t*)cnt, B, cap, pc, offs, info);
  if (int rc = check_launch("scatter_scan")) return rc;
  long long want = ((long long)rows * 32 + 255) / 256;
  const int blocks = (int)(want < 148 * 4 ? want : 148 * 4);
  launch_pdl(scatter_copy_kernel, dim3(blocks), dim3(256), 0, st, col, val, B, n_dst, per, cap, (const int32_t*)cnt,
             (const int32_t*)start, (const int32_t*)offs, pl, pv);
  return check_launch("scatter_copy");
}

extern "C" int cmmvae_slab_rows(const void* slabs, long long slab_bytes, int B, int n_src, int32_t* row_begin,
                                int32_t* row_end, void* stream) {
  CMMVAE_REQUIRE(slabs && slab_bytes % 4 == 0 && B > 0 && n_src >= 1, "slab_rows: bad arguments");
  const int rows = B * n_src;
  launch_pdl(slab_rows_kernel, dim3((rows + 255) / 256), dim3(256), 0, (cudaStream_t)stream, (const uint8_t*)slabs,
             slab_bytes, B, n_src, row_begin, row_end);
  return check_launch("slab_rows");
}
