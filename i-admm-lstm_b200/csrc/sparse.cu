// Sparse problem families: generate_data.py:96-228 stores Random_QP / Equality_QP / SVM (and
// main.py reads QPLIB / Maros-Meszaros) as scipy csc matrices and main.py:243-296 densifies them with .toarray(); every
// iteration then streams the zeros from HBM.  Here a [B, rows, n] batch is re-laid once per solve as "bitmap slabs":
//
//     mask [B][rows][S]  uint4    S = ceil(n/128); bit b of word w of slab s <-> column 128 s + 32 w + b is non-zero
//     off  [B][rows][S]  uint32   index of the slab's first value in the instance's value array
//     vals [B][cap]      float    the non-zero values of the instance, row-major (cap >= the largest instance's count)
//
// = 4 nnz + 20 rows S bytes per instance instead of 4 rows n: -36 % at the 60 % density of Random_QP's A0, -46 % at Equality_QP's
// 50 %, -95 % for the SVM family's identity blocks and the truly sparse QPLIB instances -- and, unlike CSR (8 bytes per non-zero),
// never more than the dense form plus 4 %.  The KKT passes (kkt.cu) read it with the SAME lane-to-column assignment as the dense
// form (a lane's 4 columns = one nibble of the mask), so results are bit-identical to the densified problem.
#include "common.cuh"

namespace iadmm {

constexpr int kSpThreads = 256;

static size_t sp_mask_bytes(int B, int rows, int S) { return align_up((size_t)B * rows * S * sizeof(uint4), 256); }
static size_t sp_off_bytes(int B, int rows, int S) { return align_up((size_t)B * rows * S * sizeof(uint32_t), 256); }

int sparse_view(const void* packed, int B, int rows, int n, size_t cap, SpMat* out) {
  const int S = cdiv(n, 128);
  const char* p = static_cast<const char*>(packed);
  out->mask = reinterpret_cast<const uint4*>(p);
  out->off = reinterpret_cast<const uint32_t*>(p + sp_mask_bytes(B, rows, S));
  out->vals = reinterpret_cast<const float*>(p + sp_mask_bytes(B, rows, S) + sp_off_bytes(B, rows, S));
  out->S = S;
  out->mask_stride = (size_t)rows * S;
  out->vals_stride = cap;
  return IADMM_OK;
}

// the nibble of a lane's 4 columns and the slab's four mask words (valid in every lane)
__device__ __forceinline__ uint4 slab_mask(const float* __restrict__ rowp, int n, int slab, int lane) {
  const int col = slab * 128 + lane * 4;
  uint32_t nib = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (col + e < n && rowp[col + e] != 0.0f) nib |= 1u << e;        // -0.0 == 0 compares equal: dropped like +0
  uint32_t x = nib << ((lane & 7) * 4);
  x |= __shfl_xor_sync(kFullMask, x, 1);
  x |= __shfl_xor_sync(kFullMask, x, 2);
  x |= __shfl_xor_sync(kFullMask, x, 4);                             // lanes 8w..8w+7 hold word w
  uint4 mk;
  mk.x = __shfl_sync(kFullMask, x, 0);  mk.y = __shfl_sync(kFullMask, x, 8);
  mk.z = __shfl_sync(kFullMask, x, 16); mk.w = __shfl_sync(kFullMask, x, 24);
  return mk;
}

// pass 1: masks and per-slab counts (into `off`).  One warp per (row, slab); grid = (ceil(rows*S / 8), B)
__global__ void __launch_bounds__(kSpThreads) sparse_mask_kernel(const float* __restrict__ M, int rows, int n, int S,
                                                                 uint4* __restrict__ mask, uint32_t* __restrict__ off) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const long rs = (long)blockIdx.x * (kSpThreads / 32) + (threadIdx.x >> 5);
  if (rs >= (long)rows * S) return;
  const int row = (int)(rs / S), slab = (int)(rs - (long)row * S);
  const uint4 mk = slab_mask(M + ((size_t)b * rows + row) * n, n, slab, lane);
  if (lane == 0) {
    mask[(size_t)b * rows * S + rs] = mk;
    off[(size_t)b * rows * S + rs] = __popc(mk.x) + __popc(mk.y) + __popc(mk.z) + __popc(mk.w);
  }
}

// pass 2: exclusive scan of the counts of one instance (row-major slab order), in place; total -> nnz[b]
__global__ void __launch_bounds__(1024) sparse_scan_kernel(uint32_t* __restrict__ off, long count, int* __restrict__ nnz) {
  __shared__ uint32_t warp_tot[32], warp_excl[32];
  __shared__ uint32_t carry_s, chunk_total;
  uint32_t* o = off + (size_t)blockIdx.x * count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (long base = 0; base < count; base += 1024) {
    const long i = base + tid;
    const uint32_t v = (i < count) ? o[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFullMask, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = warp_tot[lane];
      uint32_t wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, wi, d);
        if (lane >= d) wi += t;
      }
      warp_excl[lane] = wi - w;
      if (lane == 31) chunk_total = wi;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    if (i < count) o[i] = carry + warp_excl[warp] + incl - v;
    __syncthreads();
    if (tid == 0) carry_s = carry + chunk_total;
    __syncthreads();
  }
  if (tid == 0 && nnz) nnz[blockIdx.x] = (int)carry_s;
}

// pass 3: the values.  Same warp-per-(row, slab) decomposition; entries beyond `cap` are not written (the caller checks nnz)
__global__ void __launch_bounds__(kSpThreads) sparse_fill_kernel(const float* __restrict__ M, int rows, int n, int S, size_t cap,
                                                                 const uint4* __restrict__ mask, const uint32_t* __restrict__ off,
                                                                 float* __restrict__ vals) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const long rs = (long)blockIdx.x * (kSpThreads / 32) + (threadIdx.x >> 5);
  if (rs >= (long)rows * S) return;
  const int row = (int)(rs / S), slab = (int)(rs - (long)row * S);
  const uint4 mk = mask[(size_t)b * rows * S + rs];
  const uint32_t base = off[(size_t)b * rows * S + rs];
  const int w = lane >> 3, sh = (lane & 7) * 4;
  const uint32_t mine = (w == 0) ? mk.x : (w == 1) ? mk.y : (w == 2) ? mk.z : mk.w;
  uint32_t pos = base + ((w > 0) ? __popc(mk.x) : 0) + ((w > 1) ? __popc(mk.y) : 0) + ((w > 2) ? __popc(mk.z) : 0) +
                 __popc(mine & ((1u << sh) - 1u));
  const uint32_t nib = (mine >> sh) & 15u;
  const float* rowp = M + ((size_t)b * rows + row) * n + slab * 128 + lane * 4;
  float* v = vals + (size_t)b * cap;
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (nib & (1u << e)) {
      if (pos < cap) v[pos] = rowp[e];
      ++pos;
    }
}

// Block occupancy of a dense batch: bit s of word g of instance b <-> rows 8g..8g+7, columns 128s..128s+127 hold a non-zero.
// grid = (ceil(rows/8), B), 8 warps; warp w scans the slabs w, w+8, ...
__global__ void __launch_bounds__(256) block_mask_kernel(const float* __restrict__ M, int rows, int n, int S,
                                                         unsigned long long* __restrict__ blk, unsigned int* __restrict__ count) {
  __shared__ unsigned long long word_s;
  const int b = blockIdx.y, g = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) word_s = 0ull;
  __syncthreads();
  unsigned long long mine = 0ull;
  for (int s = warp; s < S; s += 8) {
    bool nz = false;
    const int col = s * 128 + lane * 4;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int row = g * 8 + u;
      if (row < rows) {
        const float* rp = M + ((size_t)b * rows + row) * n;
#pragma unroll
        for (int e = 0; e < 4; ++e) nz |= (col + e < n) && (rp[col + e] != 0.0f);
      }
    }
    if (__any_sync(kFullMask, nz)) mine |= 1ull << s;
  }
  if (lane == 0 && mine) atomicOr(&word_s, mine);
  __syncthreads();
  if (threadIdx.x == 0) {
    blk[(size_t)b * gridDim.x + g] = word_s;
    if (count) atomicAdd(count + b, (unsigned int)__popcll(word_s));
  }
}

}  // namespace iadmm

using namespace iadmm;

extern "C" {

int iadmm_block_mask_bytes(int B, int rows, int n, size_t* bytes) {
  if (B <= 0 || rows < 0 || n <= 0 || !bytes) IADMM_FAIL(IADMM_ESHAPE, "block_mask_bytes: B=%d rows=%d n=%d", B, rows, n);
  if (n > 64 * 128) IADMM_FAIL(IADMM_EMODE, "block masks cover at most 8192 columns (got %d)", n);
  *bytes = align_up((size_t)B * ((rows + 7) / 8) * sizeof(unsigned long long), 256) + 256;
  return IADMM_OK;
}

int iadmm_block_mask(const float* M, int B, int rows, int n, void* blocks, size_t blocks_bytes, unsigned int* nonempty, void* stream) {
  size_t need = 0;
  int rc = iadmm_block_mask_bytes(B, rows, n, &need);
  if (rc) return rc;
  if (B > 65535) IADMM_FAIL(IADMM_ESHAPE, "block_mask: batch %d > 65535", B);
  if (!M || !blocks || !aligned16(blocks)) IADMM_FAIL(IADMM_EALIGN, "block_mask: NULL or unaligned pointer");
  if (blocks_bytes < need) IADMM_FAIL(IADMM_EWORK, "block_mask: buffer too small: %zu < %zu", blocks_bytes, need);
  int dev = 0, major = 0;
  IADMM_CUDA(cudaGetDevice(&dev));
  IADMM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) IADMM_FAIL(IADMM_EARCH, "device %d has compute capability %d.x; libiadmm_b200 needs sm_100 (B200)", dev, major);
  if (rows == 0) return IADMM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nonempty) IADMM_CUDA(cudaMemsetAsync(nonempty, 0, (size_t)B * sizeof(unsigned int), st));
  block_mask_kernel<<<dim3((rows + 7) / 8, B), 256, 0, st>>>(M, rows, n, cdiv(n, 128), static_cast<unsigned long long*>(blocks), nonempty);
  IADMM_LAUNCH_CHECK("block_mask_kernel");
  return IADMM_OK;
}

int iadmm_sparse_bytes(int B, int rows, int n, size_t cap, size_t* bytes) {
  if (B <= 0 || rows < 0 || n <= 0 || !bytes) IADMM_FAIL(IADMM_ESHAPE, "sparse_bytes: B=%d rows=%d n=%d", B, rows, n);
  const int S = cdiv(n, 128);
  *bytes = sp_mask_bytes(B, rows, S) + sp_off_bytes(B, rows, S) + align_up((size_t)B * cap * sizeof(float), 256) + 256;
  return IADMM_OK;
}

int iadmm_sparse_pack(const float* M, int B, int rows, int n, size_t cap, void* packed, size_t packed_bytes, int* nnz,
                      void* stream) {
  if (B <= 0 || rows < 0 || n <= 0 || B > 65535) IADMM_FAIL(IADMM_ESHAPE, "sparse_pack: B=%d rows=%d n=%d", B, rows, n);
  if ((size_t)rows * n >= 0xffffffffull) IADMM_FAIL(IADMM_ESHAPE, "sparse_pack: rows*n does not fit 32-bit value offsets");
  if (!M || !packed || !aligned16(packed)) IADMM_FAIL(IADMM_EALIGN, "sparse_pack: NULL or unaligned pointer");
  size_t need = 0;
  iadmm_sparse_bytes(B, rows, n, cap, &need);
  if (packed_bytes < need) IADMM_FAIL(IADMM_EWORK, "sparse_pack: buffer too small: %zu < %zu", packed_bytes, need);
  int dev = 0, major = 0;
  IADMM_CUDA(cudaGetDevice(&dev));
  IADMM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) IADMM_FAIL(IADMM_EARCH, "device %d has compute capability %d.x; libiadmm_b200 needs sm_100 (B200)", dev, major);
  if (rows == 0) return IADMM_OK;
  SpMat v;
  sparse_view(packed, B, rows, n, cap, &v);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long count = (long)rows * v.S;
  const dim3 grid((unsigned)((count + kSpThreads / 32 - 1) / (kSpThreads / 32)), B);
  uint4* mask = const_cast<uint4*>(v.mask);
  uint32_t* off = const_cast<uint32_t*>(v.off);
  sparse_mask_kernel<<<grid, kSpThreads, 0, st>>>(M, rows, n, v.S, mask, off);
  IADMM_LAUNCH_CHECK("sparse_mask_kernel");
  sparse_scan_kernel<<<B, 1024, 0, st>>>(off, count, nnz);
  IADMM_LAUNCH_CHECK("sparse_scan_kernel");
  sparse_fill_kernel<<<grid, kSpThreads, 0, st>>>(M, rows, n, v.S, cap, mask, off, const_cast<float*>(v.vals));
  IADMM_LAUNCH_CHECK("sparse_fill_kernel");
  return IADMM_OK;
}

}  // extern "C"
