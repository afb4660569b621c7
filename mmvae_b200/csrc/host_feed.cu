// Host side of the CSR batch feed + the one device kernel it needs.
//
// cmmvae_host_slice_rows: rows [lo, hi) of a CSR chunk (what scipy's chunk[lo:hi] yields in the reference's
// batcher, cellxgene_datapipe.py:173-183) written straight into a pinned staging block: crow rebased to 0,
// gene ids narrowed to uint16 when the panel has <= 65536 genes (2 bytes less per non-zero on the PCIe wire;
// the value stays fp32, bit for bit), values copied.  Plain C loops that the host compiler vectorises; called
// through ctypes (which drops the GIL), so several batches are packed in parallel by worker threads.
// cmmvae_widen_u16_i32: the device widens the ids back to the int32 col array every kernel consumes.
#include <immintrin.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"

namespace cmmvae {
__global__ void widen_u16_i32_kernel(const uint16_t* __restrict__ src, int32_t* __restrict__ dst, long long n) {
  // 8 ids per thread: one 16-byte load, two 16-byte stores
  const long long n8 = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
    int4 a, b;
    a.x = v.x & 0xFFFF; a.y = v.x >> 16; a.z = v.y & 0xFFFF; a.w = v.y >> 16;
    b.x = v.z & 0xFFFF; b.y = v.z >> 16; b.z = v.w & 0xFFFF; b.w = v.w >> 16;
    reinterpret_cast<int4*>(dst)[2 * i] = a;
    reinterpret_cast<int4*>(dst)[2 * i + 1] = b;
  }
  for (long long i = n8 * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
__global__ void copy_bytes_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n16,
                                  const unsigned char* __restrict__ src_tail, unsigned char* __restrict__ dst_tail,
                                  int tail) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = __ldg(src + i);
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}
}  // namespace cmmvae

using namespace cmmvae;

// device -> device copy done by the SMs.  The graph-replayed step copies each staged batch to the fixed addresses its
// graphs read; as cudaMemcpyAsync those copies queue on the same copy engines as the H2D transfers of the batches
// staged ahead (0.3-0.5 ms each) and the step waits for them; as a kernel they take ~15 us and wait for nothing.
extern "C" int cmmvae_copy_bytes(void* dst, const void* src, long long nbytes, void* stream) {
  if (nbytes <= 0) return 0;
  CMMVAE_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "copy_bytes: buffers must be 16-byte aligned");
  const long long n16 = nbytes / 16;
  const int tail = (int)(nbytes - n16 * 16);
  long long want = (n16 + 255) / 256 + 1;
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
  copy_bytes_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)src, (uint4*)dst, n16, (const unsigned char*)src + n16 * 16, (unsigned char*)dst + n16 * 16, tail);
  return check_launch("copy_bytes");
}

extern "C" int cmmvae_widen_u16_i32(const void* src_u16, int32_t* dst, long long n, void* stream) {
  if (n <= 0) return 0;
  CMMVAE_REQUIRE((((uintptr_t)src_u16 | (uintptr_t)dst) & 15) == 0, "widen_u16_i32: buffers must be 16-byte aligned");
  long long want = (n / 8 + 255) / 256 + 1;
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
  widen_u16_i32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src_u16, dst, n);
  return check_launch("widen_u16_i32");
}

template <typename I, typename O>
static bool narrow_ids(const I* __restrict__ src, O* __restrict__ dst, long long n, long long n_genes) {
  // unsigned max-reduction: negative ids wrap to huge values, so one compare after the loop validates the
  // whole range and the loop body stays branch free (the host compiler vectorises it)
  typedef typename std::make_unsigned<I>::type U;
  U mx = 0;
  for (long long i = 0; i < n; ++i) {
    const U u = (U)src[i];
    mx = u > mx ? u : mx;
    dst[i] = (O)u;
  }
  return n == 0 || (unsigned long long)mx < (unsigned long long)n_genes;
}

// AVX2 route of the common case (int32 ids -> uint16, fp32 values): the pinned block is written with streaming
// stores -- it is read next by the DMA engine, not by this core, so pulling its lines into the cache first
// (read-for-ownership) would only add a third of DRAM traffic to a pass that is DRAM-bound when every rank packs
__attribute__((target("avx2"))) static bool narrow_i32_u16_avx2(const int32_t* __restrict__ src,
                                                                 uint16_t* __restrict__ dst, long long n,
                                                                 long long n_genes) {
  uint32_t mx_s = 0;
  long long i = 0;
  for (; i < n && ((uintptr_t)(dst + i) & 31); ++i) {
    const uint32_t u = (uint32_t)src[i];
    mx_s = u > mx_s ? u : mx_s;
    dst[i] = (uint16_t)u;
  }
  __m256i mx = _mm256_setzero_si256();
  for (; i + 16 <= n; i += 16) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i));
    const __m256i b = _mm256_loadu_si256((const __m256i*)(src + i + 8));
    mx = _mm256_max_epu32(mx, _mm256_max_epu32(a, b));
    // packus works per 128-bit lane: [a0-3 b0-3 | a4-7 b4-7] -> reorder the 64-bit quarters
    const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xD8);
    _mm256_stream_si256((__m256i*)(dst + i), p);
  }
  alignas(32) uint32_t m8[8];
  _mm256_store_si256((__m256i*)m8, mx);
  for (int k = 0; k < 8; ++k) mx_s = m8[k] > mx_s ? m8[k] : mx_s;
  for (; i < n; ++i) {
    const uint32_t u = (uint32_t)src[i];
    mx_s = u > mx_s ? u : mx_s;
    dst[i] = (uint16_t)u;
  }
  _mm_sfence();
  return n == 0 || (unsigned long long)mx_s < (unsigned long long)n_genes;
}

__attribute__((target("avx2"))) static void stream_copy_f32_avx2(const float* __restrict__ src,
                                                                 float* __restrict__ dst, long long n) {
  long long i = 0;
  for (; i < n && ((uintptr_t)(dst + i) & 31); ++i) dst[i] = src[i];
  for (; i + 16 <= n; i += 16) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i));
    const __m256i b = _mm256_loadu_si256((const __m256i*)(src + i + 8));
    _mm256_stream_si256((__m256i*)(dst + i), a);
    _mm256_stream_si256((__m256i*)(dst + i + 8), b);
  }
  for (; i < n; ++i) dst[i] = src[i];
  _mm_sfence();
}

static bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}

template <typename P, typename I>
static long long slice_rows_impl(const P* indptr, const I* indices, const float* data, long long lo, long long hi,
                                 int32_t* crow_out, void* col_out, int col_u16, float* val_out, long long n_genes) {
  const long long a = (long long)indptr[lo], b = (long long)indptr[hi], n = b - a;
  for (long long r = lo; r <= hi; ++r) crow_out[r - lo] = (int32_t)((long long)indptr[r] - a);
  bool ok;
  if (col_u16 && std::is_same<I, int32_t>::value && have_avx2()) {
    ok = narrow_i32_u16_avx2((const int32_t*)(indices + a), (uint16_t*)col_out, n, n_genes);
    stream_copy_f32_avx2(data + a, val_out, n);
    return ok ? n : -1;
  }
  ok = col_u16 ? narrow_ids(indices + a, (uint16_t*)col_out, n, n_genes)
               : narrow_ids(indices + a, (int32_t*)col_out, n, n_genes);
  memcpy(val_out, data + a, sizeof(float) * (size_t)n);
  return ok ? n : -1;
}

// all pointers are HOST pointers.  indptr/indices are int32 (width 4) or int64 (width 8), data is float32.
// Returns the number of non-zeros written, or a negative code (out-of-range gene id / bad arguments).
extern "C" long long cmmvae_host_slice_rows(const void* indptr, int indptr_width, const void* indices,
                                            int indices_width, const float* data, long long lo, long long hi,
                                            long long n_genes, int32_t* crow_out, void* col_out, int col_u16,
                                            float* val_out) {
  if (!indptr || !indices || !data || !crow_out || !col_out || !val_out || lo < 0 || hi < lo ||
      (col_u16 && n_genes > 65536)) {
    set_error("host_slice_rows: bad arguments");
    return -2;
  }
  long long n;
  if (indptr_width == 4 && indices_width == 4)
    n = slice_rows_impl((const int32_t*)indptr, (const int32_t*)indices, data, lo, hi, crow_out, col_out, col_u16, val_out, n_genes);
  else if (indptr_width == 8 && indices_width == 8)
    n = slice_rows_impl((const int64_t*)indptr, (const int64_t*)indices, data, lo, hi, crow_out, col_out, col_u16, val_out, n_genes);
  else if (indptr_width == 8 && indices_width == 4)
    n = slice_rows_impl((const int64_t*)indptr, (const int32_t*)indices, data, lo, hi, crow_out, col_out, col_u16, val_out, n_genes);
  else if (indptr_width == 4 && indices_width == 8)
    n = slice_rows_impl((const int32_t*)indptr, (const int64_t*)indices, data, lo, hi, crow_out, col_out, col_u16, val_out, n_genes);
  else {
    set_error("host_slice_rows: index width must be 4 or 8 bytes");
    return -2;
  }
  if (n < 0) set_error("host_slice_rows: gene id outside [0, %lld)", n_genes);
  return n;
}

// ---- page-locked chunks: batches are DMA'd straight out of the chunk's own arrays ------------------------------
extern "C" int cmmvae_host_register(const void* host_ptr, long long nbytes) {
  CMMVAE_REQUIRE(host_ptr && nbytes > 0, "host_register: bad arguments");
  cudaError_t e = cudaHostRegister(const_cast<void*>(host_ptr), (size_t)nbytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("host_register(%lld bytes): %s", nbytes, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

extern "C" int cmmvae_host_unregister(const void* host_ptr) {
  cudaError_t e = cudaHostUnregister(const_cast<void*>(host_ptr));
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("host_unregister: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

extern "C" int cmmvae_h2d_async(void* dst_dev, const void* src_host, long long nbytes, void* stream) {
  if (nbytes <= 0) return 0;
  CMMVAE_REQUIRE(dst_dev && src_host, "h2d_async: null pointer");
  cudaError_t e = cudaMemcpyAsync(dst_dev, src_host, (size_t)nbytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("h2d_async(%lld bytes): %s", nbytes, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}
