// Weight gradients of the backward pass on the tensor cores:  dW[n][k] = sum_m dY[m][n] X[m][k],  db[n] = sum_m dY[m][n]
// -- what torch.autograd derives for every nn.Linear / Conv1d(k=1) of renderer.py:350,423-424,920-993 and triplane.py:285-316.
//
// It is a GEMM whose REDUCTION runs over the surviving points: D[n][k] (128 x <=208, fp32 in TMEM) += A[n][p] . B[k][p]^T with p the
// tcgen05 K dimension.  dY / X sit in HBM row-major over points, so a 32-point chunk is transposed on its way into shared memory: each
// thread loads a 4-point x 4-feature block (four 16-byte loads), splits it into tf32 hi / lo and stores four 16-byte core-matrix rows
// (feature-major) of the K-major no-swizzle canonical layout of umma.cuh.  The 8 lanes of a quarter-warp write the 8 point-quads of one
// feature block: with the K-direction stride padded by 16 bytes their STS.128 are bank-conflict free, and a warp's loads cover 8 rows x 64
// contiguous bytes.  Arithmetic: 3xTF32 split products (a_lo w_hi + a_hi w_lo + a_hi w_hi, fp32 accumulate) = fp32-grade sums.
// The bias gradient (column sums of dY) is accumulated in fp32 registers by the threads that stage dY and written as column K of the
// partial tile.  The launcher deals the 32-point chunks to one wave of CTAs in equal contiguous ranges (no tail wave); every CTA writes its
// partial [n][K + 1] tile and k_reduce_parts (backward.cu) adds the partials in split order: the result does not depend on scheduling.
#include "common.cuh"
#include "stages.cuh"
#include "umma.cuh"

namespace sherf {

struct GradWArgs {
  const float* dY; int lda, agroup, agstride; int N;     // [M][N], logical column c at (c / agroup) * agstride + c % agroup
  const float* X; int ldb; int K;                         // [M][K]
  float* part;                                            // [splits][N][K + 1]
  int M, rows_per_split;
  int Np;                                                 // round_up(K, 16)
  uint32_t tmem_cols;
};

__device__ __forceinline__ void store_split(unsigned char* hi, unsigned char* lo, uint32_t off, float a, float b, float c, float d) {
  const float4 h = make_float4(umma::to_tf32(a), umma::to_tf32(b), umma::to_tf32(c), umma::to_tf32(d));
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = make_float4(umma::to_tf32(a - h.x), umma::to_tf32(b - h.y), umma::to_tf32(c - h.z), umma::to_tf32(d - h.w));
}

constexpr int kGwChunk = 32;                     // points per shared-memory chunk = 8 core-matrix columns = 4 MMA k-steps
constexpr uint32_t kGwALbo = 128 * 16 + 16;      // K-direction stride of the A operand (128 feature rows), padded

// NU = feature blocks of X per thread: 1 for K <= 128 (three CTAs per SM), 2 up to K = 256
template <int NU>
__global__ void __launch_bounds__(256, NU == 1 ? 3 : 2) k_umma_grad_w(const GradWArgs g) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * 128, K1 = g.K + 1;
  const int mlo = blockIdx.x * g.rows_per_split, mhi = min(g.M, mlo + g.rows_per_split);
  const uint32_t b_lbo = (uint32_t)g.Np * 16u + 16u;
  const uint32_t a_bytes = 8 * kGwALbo, b_bytes = 8 * b_lbo;
  unsigned char* A_hi = smem;
  unsigned char* A_lo = smem + a_bytes;
  unsigned char* B_hi = smem + 2 * a_bytes;
  unsigned char* B_lo = smem + 2 * a_bytes + b_bytes;

  if (tid == 0) { umma::mbar_init(&mma_bar, 1); umma::fence_mbar_init(); }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, g.tmem_cols);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = umma::make_idesc_tf32(128, g.Np);

  const int p4 = tid & 7;                        // point quad of the chunk this thread transposes
  const int fb0 = tid >> 3;                      // first feature block (4 features); A has 32 of them, B has Np / 4 <= 64
  const int nbB = g.Np / 4;
  // one chunk = a 4-point x 4-feature block of dY and up to two of [X | 1] per thread, held in registers between the loads (issued while
  // the previous chunk's MMAs run) and the transposing stores.  The loads are branch-free in the common case and NOTHING consumes them
  // before store_chunk (zero fill of rows past the split and the ones column are applied there): twelve 16-byte loads per thread stay in
  // flight together (ncu on the first version: 48 % of the stall samples sat on a load whose value was patched right behind it).
  // Per unit, decided once: mode 0 = all zeros / ones column only, 1 = one 16-byte load per row, 2 = guarded scalar loads (row tails).
  float4 va[4], vb[NU][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};            // column sums of this thread's dY block over all its chunks (bias gradient)
  const int ca = n0 + fb0 * 4;
  const float* pa = g.dY + (g.agroup ? (ca / g.agroup) * g.agstride + (ca % g.agroup) : ca);
  const int mode_a = ca >= g.N ? 0 : ((ca + 3 < g.N && (g.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(pa) & 15) == 0) ? 1 : 2);
  const float* pb[NU];
  int mode_b[NU], cb[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int fb = fb0 + 32 * u;
    cb[u] = fb * 4;
    pb[u] = g.X + cb[u];
    mode_b[u] = (fb >= nbB || cb[u] >= g.K) ? 0 : ((cb[u] + 3 < g.K && (g.ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(pb[u]) & 15) == 0) ? 1 : 2);
  }
  auto load_rows = [&](float4 (&v)[4], const float* p, int ld, int mode, int c, int ncols, int mrow) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = min(mrow + j, g.M - 1);                                // clamped: always a valid address; masked in store_chunk
      const float* q = p + (size_t)row * ld;
      if (mode == 1) v[j] = __ldg(reinterpret_cast<const float4*>(q));
      else if (mode == 2) {
        v[j].x = __ldg(q);
        v[j].y = c + 1 < ncols ? __ldg(q + 1) : 0.f;
        v[j].z = c + 2 < ncols ? __ldg(q + 2) : 0.f;
        v[j].w = c + 3 < ncols ? __ldg(q + 3) : 0.f;
      } else v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto load_chunk = [&](int mb) {
    const int mrow = mb + p4 * 4;
    load_rows(va, pa, g.lda, mode_a, ca, g.N, mrow);
#pragma unroll
    for (int u = 0; u < NU; ++u) load_rows(vb[u], pb[u], g.ldb, mode_b[u], cb[u], g.K, mrow);
  };
  auto store_chunk = [&](int mb) {
    const int mrow = mb + p4 * 4;
    float rowok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) rowok[j] = mrow + j < mhi ? 1.f : 0.f;
    {
      const uint32_t off = (uint32_t)p4 * kGwALbo + (uint32_t)(fb0 * 4) * 16u;
      store_split(A_hi, A_lo, off, va[0].x * rowok[0], va[1].x * rowok[1], va[2].x * rowok[2], va[3].x * rowok[3]);
      store_split(A_hi, A_lo, off + 16, va[0].y * rowok[0], va[1].y * rowok[1], va[2].y * rowok[2], va[3].y * rowok[3]);
      store_split(A_hi, A_lo, off + 32, va[0].z * rowok[0], va[1].z * rowok[1], va[2].z * rowok[2], va[3].z * rowok[3]);
      store_split(A_hi, A_lo, off + 48, va[0].w * rowok[0], va[1].w * rowok[1], va[2].w * rowok[2], va[3].w * rowok[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bsum[0] = fmaf(va[j].x, rowok[j], bsum[0]); bsum[1] = fmaf(va[j].y, rowok[j], bsum[1]);
        bsum[2] = fmaf(va[j].z, rowok[j], bsum[2]); bsum[3] = fmaf(va[j].w, rowok[j], bsum[3]);
      }
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int fb = fb0 + 32 * u;
      if (fb < nbB) {
        const uint32_t off = (uint32_t)p4 * b_lbo + (uint32_t)(fb * 4) * 16u;
        store_split(B_hi, B_lo, off, vb[u][0].x * rowok[0], vb[u][1].x * rowok[1], vb[u][2].x * rowok[2], vb[u][3].x * rowok[3]);
        store_split(B_hi, B_lo, off + 16, vb[u][0].y * rowok[0], vb[u][1].y * rowok[1], vb[u][2].y * rowok[2], vb[u][3].y * rowok[3]);
        store_split(B_hi, B_lo, off + 32, vb[u][0].z * rowok[0], vb[u][1].z * rowok[1], vb[u][2].z * rowok[2], vb[u][3].z * rowok[3]);
        store_split(B_hi, B_lo, off + 48, vb[u][0].w * rowok[0], vb[u][1].w * rowok[1], vb[u][2].w * rowok[2], vb[u][3].w * rowok[3]);
      }
    }
  };
  uint32_t parity = 0;
  bool first = true;
  if (mlo < mhi) load_chunk(mlo);
  for (int mb = mlo; mb < mhi; mb += kGwChunk) {
    if (!first) { umma::mbar_wait(&mma_bar, parity); parity ^= 1; }       // the previous chunk's MMAs have read the operands
    store_chunk(mb);
    umma::fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      umma::tc_fence_after_sync();
      const uint32_t a_hi_s = umma::smem_u32(A_hi), a_lo_s = umma::smem_u32(A_lo), b_hi_s = umma::smem_u32(B_hi), b_lo_s = umma::smem_u32(B_lo);
#pragma unroll
      for (int s = 0; s < 4; ++s) {                                         // MMA K = 8 points = two core-matrix columns
        const uint32_t a_off = (uint32_t)s * 2u * kGwALbo, b_off = (uint32_t)s * 2u * b_lbo;
        const uint64_t ah = umma::make_smem_desc(a_hi_s + a_off, kGwALbo, 128u), al = umma::make_smem_desc(a_lo_s + a_off, kGwALbo, 128u);
        const uint64_t bh = umma::make_smem_desc(b_hi_s + b_off, b_lbo, 128u), bl = umma::make_smem_desc(b_lo_s + b_off, b_lbo, 128u);
        umma::mma_tf32_ss_w(tmem_base, al, bh, idesc, (first && s == 0) ? 0u : 1u);      // small terms first
        umma::mma_tf32_ss_w(tmem_base, ah, bl, idesc, 1u);
        umma::mma_tf32_ss_w(tmem_base, ah, bh, idesc, 1u);
      }
      umma::mma_commit_w(&mma_bar);
      __syncwarp();
    }
    first = false;
    if (mb + kGwChunk < mhi) load_chunk(mb + kGwChunk);                     // in flight while the tensor cores work on this chunk
  }
  if (!first) {
    umma::mbar_wait(&mma_bar, parity);
    umma::tc_fence_after_sync();
    // ---- epilogue: TMEM lane = output feature n, column = input feature k.  Warp w owns lane quarter (w & 3), column half (w >> 2). ----
    const int q = warp & 3, hsel = warp >> 2;
    const int n = n0 + 32 * q + lane;
    const int half = g.Np / 2;                                             // Np % 16 == 0: a multiple of 8
    float* out = g.part + ((size_t)blockIdx.x * g.N + n) * K1;
    for (int j = 0; j < half / 8; ++j) {
      const int c0 = hsel * half + 8 * j;
      uint32_t v[8];
      umma::tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
      umma::tmem_ld_wait();
      if (n < g.N) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c0 + i < g.K) out[c0 + i] = __uint_as_float(v[i]);
      }
    }
    // bias column: add the eight point-quad lanes of a feature block (xor 1, 2, 4 stay inside the quarter-warp), lane p4 == 0 writes
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float b = bsum[i];
      b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2); b += __shfl_xor_sync(0xffffffffu, b, 4);
      const int nn = n0 + fb0 * 4 + i;
      if (p4 == 0 && nn < g.N) g.part[((size_t)blockIdx.x * g.N + nn) * K1 + g.K] = b;
    }
  } else if (mlo >= mhi) {
    // empty split (cannot happen with the launcher's grid, kept for safety): its partial tile must still be defined
    for (int i = tid; i < min(128, g.N - n0) * K1; i += 256) g.part[((size_t)blockIdx.x * g.N + n0) * K1 + i] = 0.f;
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, g.tmem_cols);
}

int launch_umma_grad_w(const float* dY, int lda, int N, const float* X, int ldb, int K, int M, float* part, int* splits_out, cudaStream_t st,
                       int agroup, int agstride) {
  if (splits_out) *splits_out = 0;
  if (M <= 0) return SHERF_OK;
  GradWArgs g;
  g.dY = dY; g.lda = lda; g.agroup = agroup; g.agstride = agstride; g.N = N; g.X = X; g.ldb = ldb; g.K = K; g.part = part; g.M = M;
  g.Np = (K + 15) / 16 * 16;
  if (g.Np > 256) { set_error("grad_w: K = %d exceeds one MMA tile", K); return SHERF_E_INVALID; }
  uint32_t cols = 32;
  while ((int)cols < g.Np) cols <<= 1;
  g.tmem_cols = cols;
  const size_t smem = 2 * (size_t)8 * kGwALbo + 2 * (size_t)8 * ((size_t)g.Np * 16 + 16);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    SHERF_CUDA_OK(cudaGetDevice(&dev));
    SHERF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_umma_grad_w<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_umma_grad_w<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  // one wave: the 32-point chunks are dealt to (SMs x resident CTAs) splits in equal contiguous ranges.  Residency: registers admit three
  // CTAs of the one-block variant (TMEM: 3 x <= 128 columns) and two of the two-block variant (2 x 256 columns): tcgen05.alloc never waits.
  const int nu = g.Np > 128 ? 2 : 1;
  const int ntiles = ceil_div(N, 128);
  const int slots = max(1, sms * (nu == 1 ? 3 : 2) / ntiles);
  const int nchunks = ceil_div(M, kGwChunk);
  const int min_rows = ceil_div(ceil_div(M, kGradWMaxSplits), kGwChunk) * kGwChunk;       // never more partial tiles than the buffer holds
  g.rows_per_split = max(ceil_div(nchunks, slots) * kGwChunk, min_rows);
  const int splits = ceil_div(M, g.rows_per_split);
  if (splits_out) *splits_out = splits;
  if (nu == 1) k_umma_grad_w<1><<<dim3(splits, ntiles), 256, smem, st>>>(g);
  else k_umma_grad_w<2><<<dim3(splits, ntiles), 256, smem, st>>>(g);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

}  // namespace sherf
