// Backward pass of the hot path: what torch.autograd derives for ImportanceRenderer.forward (renderer.py:286-398),
// run_model (:400-437), NeRFDecoder.forward (triplane.py:285-316) and MipRayMarcher2 (ray_marcher.py:25-64) when the reference calls
// loss.backward() (loss.py:175).  Recompute-in-backward: nothing is kept from the forward but its per-point sigma / rgb; every chunk of
// surviving points is gathered again, pushed through the fp32 per-layer path with all activations kept, and walked back layer by layer.
//
//   dL/d(rgb, depth, acc)[N] --k_composite_bwd--> dL/d(sigma, rgb)[P]
//   per chunk:  rgb head, views, feature, alpha, pts 7..0  (dX = dY W, dW += dY^T X, db += sum dY)
//               -> d tok0 / d tok1 -> FeedForward, LayerNorm-2, to_out, 3-token attention, to_qkv, LayerNorm-1 -> conv1d_reprojection
//               -> conv1d_projection -> adjoint of the three gathers (gather.cu, vector reductions into channels-last gradient grids)
//
// Coordinates carry no gradient: the warps depend only on SMPL parameters and cameras, which are data (requires_grad False upstream).
// Weight gradients are reduced deterministically (per-split partial sums, added in split order); the grid gradients use red.add like
// F.grid_sample's own backward.  All arithmetic fp32 FMA.
#include "common.cuh"
#include "stages.cuh"

namespace sherf {

// ------------------------------------------------------------------------------------------------- ray marcher
// thread per ray.  w_i = alpha_i T_i, T_i = prod_{j<i} f_j, f_j = 1 - alpha_j + 1e-10                      ray_marcher.py:39-50
//   dL/dw_i = 2 g_rgb . c_i (- 2 sum g_rgb if white_back) + g_acc + g_depth (t_i - depth) / wsum
//   dL/dalpha_i = dL/dw_i T_i - (sum_{j>i} dL/dw_j w_j) / f_i ;  dalpha/dsigma = delta exp(-sigma delta) [sigma > 0]
// dsig / drgb double as scratch for w_i / T_i between the forward and the reverse sweep.
__global__ void __launch_bounds__(128) k_composite_bwd(const float* __restrict__ dirs, const float* __restrict__ nearv, const float* __restrict__ farv,
                                                       int N, int S, const FrameConst* __restrict__ fc, const int* __restrict__ ray_start,
                                                       const int* __restrict__ point_sample, const float* __restrict__ sigma,
                                                       const float* __restrict__ rgb, const float* __restrict__ noise, int white_back,
                                                       const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
                                                       const float* __restrict__ g_acc, float* __restrict__ dsig, float* __restrict__ drgb) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int b = ray_start[n], e = ray_start[n + 1];
  if (b >= e) return;
  const float dx = dirs[n * 3], dy = dirs[n * 3 + 1], dz = dirs[n * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float nr = nearv[n], fr = farv[n];
  float T = 1.f, wsum = 0.f, wdepth = 0.f;
  for (int p = b; p < e; ++p) {
    const int i = point_sample[p] - n * S;
    const float t = sample_depth(nr, fr, i, S);
    const float delta = ((i == S - 1) ? 1e10f : (sample_depth(nr, fr, i + 1, S) - t)) * dnorm;
    float sg = sigma[p];
    if (noise) sg += noise[p];
    const float alpha = 1.f - expf(-(fmaxf(sg, 0.f) * delta));
    const float w = alpha * T;
    dsig[p] = w;
    drgb[(size_t)p * 3] = T;
    wsum += w; wdepth += w * t;
    T *= (1.f - alpha + 1e-10f);
  }
  const float gr = g_rgb ? g_rgb[(size_t)n * 3] : 0.f, gg = g_rgb ? g_rgb[(size_t)n * 3 + 1] : 0.f, gb = g_rgb ? g_rgb[(size_t)n * 3 + 2] : 0.f;
  const float ga = g_acc ? g_acc[n] : 0.f;
  const float depth = wdepth / wsum;
  // torch.clamp passes the gradient inside [min, max]; 0/0 -> nan -> +inf -> clamped: no gradient                ray_marcher.py:53-57
  const bool depth_live = g_depth && depth == depth && depth >= ordered_to_float(fc->dmin_bits) && depth <= ordered_to_float(fc->dmax_bits);
  const float gd = depth_live ? g_depth[n] / wsum : 0.f;
  const float gwb = white_back ? -2.f * (gr + gg + gb) : 0.f;
  float suffix = 0.f;
  for (int p = e - 1; p >= b; --p) {
    const int i = point_sample[p] - n * S;
    const float t = sample_depth(nr, fr, i, S);
    const float delta = ((i == S - 1) ? 1e10f : (sample_depth(nr, fr, i + 1, S) - t)) * dnorm;
    float sg = sigma[p];
    if (noise) sg += noise[p];
    const float ex = expf(-(fmaxf(sg, 0.f) * delta));
    const float f = (1.f - (1.f - ex)) + 1e-10f;
    const float w = dsig[p], Tp = drgb[(size_t)p * 3];
    const float c0 = rgb[(size_t)p * 3], c1 = rgb[(size_t)p * 3 + 1], c2 = rgb[(size_t)p * 3 + 2];
    const float G = 2.f * (gr * c0 + gg * c1 + gb * c2) + gwb + ga + gd * (t - depth);
    const float dalpha = G * Tp - suffix / f;
    suffix += G * w;
    dsig[p] = sg > 0.f ? dalpha * (delta * ex) : 0.f;
    drgb[(size_t)p * 3] = 2.f * w * gr; drgb[(size_t)p * 3 + 1] = 2.f * w * gg; drgb[(size_t)p * 3 + 2] = 2.f * w * gb;
  }
}

int run_composite_backward(const SherfRays& rays, const FrameConst* fc, const int* ray_start, const int* point_sample, const float* sigma,
                           const float* rgb, const float* noise, int white_back, const float* g_rgb, const float* g_depth, const float* g_acc,
                           float* dsig, float* drgb, cudaStream_t st) {
  k_composite_bwd<<<ceil_div(rays.n_rays, 128), 128, 0, st>>>(rays.dirs, rays.near_, rays.far_, rays.n_rays, rays.n_samples, fc, ray_start,
                                                              point_sample, sigma, rgb, noise, white_back, g_rgb, g_depth, g_acc, dsig, drgb);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

#define RC(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

// ------------------------------------------------------------------------------------------------- GEMMs
// column map of an operand: logical column c -> (c / group) * gstride + c % group   (group = 0: identity); used for the 96 outputs of
// conv1d_projection, which live at columns 64..95 of each token's 96-wide slice of comb (renderer.py:350,423)
__device__ __forceinline__ int colmap(int c, int group, int gstride) { return group ? (c / group) * gstride + (c % group) : c; }

// C[M][Nc] (+)= A[M][Kr] . B[Kr][Nc]  (dX = dY . W with W in PyTorch's [out][in] layout: no packing).  128 x 64 tile, 8 x 4 per thread.
struct NnArgs {
  const float* A; int lda, agroup, agstride;
  const float* B; int ldb;
  float* C; int ldc;
  const float* Mask; int ldm;          // optional: result forced to 0 where Mask <= 0 (ReLU of the layer that produced this input)
  int accum;                           // C += instead of C =
  int M, Nc, Kr;
};

__global__ void __launch_bounds__(256) k_gemm_nn(const NnArgs g) {
  constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int arow = tid & 127, akh = tid >> 7;                 // A loader: row, k-half (8 consecutive k)
  const bool arow_ok = (m0 + arow) < g.M;
  const float* aptr = g.A + (size_t)(m0 + arow) * g.lda;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float ra[8], rb[4];
  const int ktiles = (g.Kr + BK - 1) / BK;
  auto load_tile = [&](int kt) {
    const int k0 = kt * BK;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int k = k0 + akh * 8 + h;
      ra[h] = (arow_ok && k < g.Kr) ? aptr[colmap(k, g.agroup, g.agstride)] : 0.f;
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int idx = tid + h * 256, k = idx >> 6, n = idx & 63;
      rb[h] = (k0 + k < g.Kr && n0 + n < g.Nc) ? g.B[(size_t)(k0 + k) * g.ldb + n0 + n] : 0.f;
    }
  };
  load_tile(0);
  for (int kt = 0; kt < ktiles; ++kt) {
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 8; ++h) As[akh * 8 + h][arow] = ra[h];
#pragma unroll
    for (int h = 0; h < 4; ++h) { const int idx = tid + h * 256; Bs[idx >> 6][idx & 63] = rb[h]; }
    __syncthreads();
    if (kt + 1 < ktiles) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.Nc) continue;
      float v = acc[i][j];
      float* c = g.C + (size_t)m * g.ldc + n;
      if (g.accum) v += *c;
      if (g.Mask && !(g.Mask[(size_t)m * g.ldm + n] > 0.f)) v = 0.f;
      *c = v;
    }
  }
}

static int gemm_nn(const float* A, int lda, const float* B, int ldb, float* Cc, int ldc, int M, int Nc, int Kr, cudaStream_t st,
                   const float* Mask = nullptr, int ldm = 0, int accum = 0, int agroup = 0, int agstride = 0) {
  NnArgs g;
  g.A = A; g.lda = lda; g.agroup = agroup; g.agstride = agstride; g.B = B; g.ldb = ldb; g.C = Cc; g.ldc = ldc; g.Mask = Mask; g.ldm = ldm;
  g.accum = accum; g.M = M; g.Nc = Nc; g.Kr = Kr;
  k_gemm_nn<<<dim3(ceil_div(M, 128), ceil_div(Nc, 64)), 256, 0, st>>>(g);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

// part[s][n][k] = sum over the rows m of split s of A[m][n] . Bx[m][k], Bx = [B | 1]  (k == K is the bias column): dW and db in one pass.
// 128 (n) x 64 (k) tile, 8 x 4 per thread, 16 rows per step.
struct TnArgs {
  const float* A; int lda, agroup, agstride; int N;
  const float* B; int ldb; int K;
  float* part; int M; int rows_per_split;
};

__global__ void __launch_bounds__(256) k_gemm_tn(const TnArgs g) {
  constexpr int BN = 128, BKc = 64, BMr = 16, TN = 8, TK = 4;
  __shared__ __align__(16) float As[BMr][BN];
  __shared__ __align__(16) float Bs[BMr][BKc];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN, k0 = blockIdx.y * BKc, K1 = g.K + 1;
  const int mlo = blockIdx.z * g.rows_per_split, mhi = min(g.M, mlo + g.rows_per_split);
  float acc[TN][TK];
#pragma unroll
  for (int i = 0; i < TN; ++i)
#pragma unroll
    for (int j = 0; j < TK; ++j) acc[i][j] = 0.f;
  float ra[8], rb[4];
  auto load_tile = [&](int mb) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int idx = tid + h * 256, mm = idx >> 7, n = idx & 127;
      const int m = mb + mm;
      ra[h] = (m < mhi && n0 + n < g.N) ? g.A[(size_t)m * g.lda + colmap(n0 + n, g.agroup, g.agstride)] : 0.f;
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int idx = tid + h * 256, mm = idx >> 6, k = idx & 63;
      const int m = mb + mm, kk = k0 + k;
      rb[h] = (m < mhi && kk < K1) ? (kk < g.K ? g.B[(size_t)m * g.ldb + kk] : 1.f) : 0.f;
    }
  };
  if (mlo < mhi) load_tile(mlo);
  for (int mb = mlo; mb < mhi; mb += BMr) {
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 8; ++h) { const int idx = tid + h * 256; As[idx >> 7][idx & 127] = ra[h]; }
#pragma unroll
    for (int h = 0; h < 4; ++h) { const int idx = tid + h * 256; Bs[idx >> 6][idx & 63] = rb[h]; }
    __syncthreads();
    if (mb + BMr < mhi) load_tile(mb + BMr);
#pragma unroll
    for (int mm = 0; mm < BMr; ++mm) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[mm][ty * TN]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[mm][ty * TN + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[mm][tx * TK]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < TN; ++i)
#pragma unroll
        for (int j = 0; j < TK; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* out = g.part + (size_t)blockIdx.z * g.N * K1;
#pragma unroll
  for (int i = 0; i < TN; ++i) {
    const int n = n0 + ty * TN + i;
    if (n >= g.N) continue;
#pragma unroll
    for (int j = 0; j < TK; ++j) {
      const int k = k0 + tx * TK + j;
      if (k < K1) out[(size_t)n * K1 + k] = acc[i][j];
    }
  }
}

// dW[n][k] += sum_s part[s][n][k]  (k < K),  db[n] += sum_s part[s][n][K].  32 outputs per block; warp w adds the splits w, w + 8, ... in
// order (128 contiguous bytes per split), the eight warp sums are added in warp order: the result does not depend on scheduling
__global__ void __launch_bounds__(256) k_reduce_parts(const float* __restrict__ part, int splits, int N, int K, float* __restrict__ dW, float* __restrict__ db) {
  __shared__ float red[8][32];
  const int K1 = K + 1, total = N * K1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (idx < total)
    for (int sp = w; sp < splits; sp += 8) s += part[(size_t)sp * total + idx];
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && idx < total) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    const int n = idx / K1, k = idx - n * K1;
    if (k < K) dW[(size_t)n * K + k] += t;
    else if (db) db[n] += t;
  }
}

constexpr int kRowsPerSplit = 1024;          // rows per split of the fp32 FMA anchor path (SHERF_BWD_SIMT=1)

// SHERF_BWD_SIMT=1: every product of the backward on the CUDA cores in fp32 FMA (the first version; kept as the anchor of the tensor-core path)
static bool bwd_simt() {
  const char* e = getenv("SHERF_BWD_SIMT");        // read per call: the tests switch it between two backward passes of one process
  return e && e[0] == '1';
}

// dW[N][K] += dY^T X, db[N] += column sums of dY     (dY = A [M][N], X = B [M][K])
static int grad_w(const float* dY, int lda, int N, const float* X, int ldb, int K, int M, float* dW, float* db, float* part, cudaStream_t st,
                  int agroup = 0, int agstride = 0) {
  if (!dW && !db) return SHERF_OK;
  if (!bwd_simt()) {
    int splits = 0;
    RC(launch_umma_grad_w(dY, lda, N, X, ldb, K, M, part, &splits, st, agroup, agstride));
    k_reduce_parts<<<ceil_div(N * (K + 1), 32), 256, 0, st>>>(part, splits, N, K, dW, db);
    SHERF_LAUNCH_CHECK();
    return SHERF_OK;
  }
  TnArgs g;
  g.A = dY; g.lda = lda; g.agroup = agroup; g.agstride = agstride; g.N = N; g.B = X; g.ldb = ldb; g.K = K; g.part = part; g.M = M;
  g.rows_per_split = kRowsPerSplit;
  const int splits = ceil_div(M, kRowsPerSplit);
  k_gemm_tn<<<dim3(ceil_div(N, 128), ceil_div(K + 1, 64), splits), 256, 0, st>>>(g);
  SHERF_LAUNCH_CHECK();
  k_reduce_parts<<<ceil_div(N * (K + 1), 32), 256, 0, st>>>(part, splits, N, K, dW, db);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

// ------------------------------------------------------------------------------------------------- element-wise pieces
__global__ void k_gelu_fwd(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float v = x[i]; y[i] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
}
// d gelu(x) = Phi(x) + x phi(x)                                                            nn.GELU() (renderer.py:953), exact erf form
__global__ void k_gelu_bwd(float* __restrict__ dy, const float* __restrict__ x, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = x[i];
    dy[i] *= 0.5f * (1.f + erff(v * 0.70710678118654752440f)) + v * 0.3989422804014327f * expf(-0.5f * v * v);
  }
}

// rgb = sigmoid(vh Wrgb^T + b) * 1.002 - 0.001 (triplane.py:312-315): dpre = drgb * 1.002 * s (1 - s); dvh = (vh > 0) dpre . Wrgb
__global__ void __launch_bounds__(256) k_head_bwd(const float* __restrict__ drgb, const float* __restrict__ rgb, const float* __restrict__ vh,
                                                  const float* __restrict__ wrgb, float* __restrict__ dpre, float* __restrict__ dvh, int np) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = idx >> 6, j = idx & 63;
  if (p >= np) return;
  float d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s = (rgb[(size_t)p * 3 + c] + 0.001f) * (1.f / 1.002f);
    d[c] = drgb[(size_t)p * 3 + c] * 1.002f * s * (1.f - s);
  }
  if (j < 4) dpre[(size_t)p * 4 + j] = j < 3 ? d[j] : 0.f;
  const float v = vh[(size_t)p * 64 + j];
  dvh[(size_t)p * 64 + j] = v > 0.f ? d[0] * wrgb[j] + d[1] * wrgb[64 + j] + d[2] * wrgb[128 + j] : 0.f;
}

// dh7 = (h7 > 0) (dfeat . Wf + dsigma alpha_w)        triplane.py:303-306
__global__ void __launch_bounds__(256) k_dh7(float* __restrict__ dh, const float* __restrict__ h7, const float* __restrict__ dsig,
                                             const float* __restrict__ aw, int np) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = idx >> 7, j = idx & 127;
  if (p >= np) return;
  const float v = dh[idx] + dsig[p] * aw[j];
  dh[idx] = h7[idx] > 0.f ? v : 0.f;
}

// dtok3[p][0] = dx[p][39:71], dtok3[p][1] = dfv[p][155:187], dtok3[p][2] = 0      (renderer.py:432, triplane.py:293,308)
__global__ void __launch_bounds__(256) k_tok3_grad(const float* __restrict__ dx, const float* __restrict__ dfv, float* __restrict__ dtok3, int np) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = idx / 96, r = idx - p * 96;
  if (p >= np) return;
  const int t = r >> 5, c = r & 31;
  dtok3[idx] = t == 0 ? dx[(size_t)p * 72 + 39 + c] : (t == 1 ? dfv[(size_t)p * 188 + 155 + c] : 0.f);
}

// LayerNorm(32) backward, eight lanes per row (16 bytes each, four rows per warp pass):
//   dx = res + rstd (g - mean(g) - xhat mean(g xhat)), g = dy w   (res = the residual branch's gradient: x + f(LN(x)), renderer.py:980-993)
// and per-block partial sums of dw = dy xhat, db = dy (added in block order by k_ln_reduce).
__global__ void __launch_bounds__(256) k_ln_bwd(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ dy,
                                                const float* __restrict__ res, float* __restrict__ dx, int rows, float* __restrict__ partial) {
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane >> 3, c4 = lane & 7;
  const float4 w4 = *reinterpret_cast<const float4*>(w + c4 * 4);
  float sw[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
  auto sum8 = [](float v) { v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4); return v; };
  for (int rb = (blockIdx.x * 8 + wid) * 4; rb < rows; rb += gridDim.x * 32) {      // warp-uniform trip count: the shuffles stay converged
    const int r = rb + sub;
    const bool ok = r < rows;
    const size_t o = (size_t)(ok ? r : 0) * 32 + c4 * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + o);
    float4 gy = *reinterpret_cast<const float4*>(dy + o);
    if (!ok) gy = make_float4(0.f, 0.f, 0.f, 0.f);
    const float mean = sum8((v.x + v.y) + (v.z + v.w)) * (1.f / 32.f);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    const float rstd = rsqrtf(sum8((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) * (1.f / 32.f) + 1e-5f);
    const float h0 = d0 * rstd, h1 = d1 * rstd, h2 = d2 * rstd, h3 = d3 * rstd;
    sw[0] += gy.x * h0; sw[1] += gy.y * h1; sw[2] += gy.z * h2; sw[3] += gy.w * h3;
    sb[0] += gy.x; sb[1] += gy.y; sb[2] += gy.z; sb[3] += gy.w;
    const float g0 = gy.x * w4.x, g1 = gy.y * w4.y, g2 = gy.z * w4.z, g3 = gy.w * w4.w;
    const float m1 = sum8((g0 + g1) + (g2 + g3)) * (1.f / 32.f);
    const float m2 = sum8((g0 * h0 + g1 * h1) + (g2 * h2 + g3 * h3)) * (1.f / 32.f);
    if (ok) {
      float4 out = res ? *reinterpret_cast<const float4*>(res + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      out.x += rstd * (g0 - m1 - h0 * m2); out.y += rstd * (g1 - m1 - h1 * m2);
      out.z += rstd * (g2 - m1 - h2 * m2); out.w += rstd * (g3 - m1 - h3 * m2);
      *reinterpret_cast<float4*>(dx + o) = out;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {                              // the four rows of a warp pass share a channel quad: add them (xor 8, 16)
    sw[i] += __shfl_xor_sync(0xffffffffu, sw[i], 8); sw[i] += __shfl_xor_sync(0xffffffffu, sw[i], 16);
    sb[i] += __shfl_xor_sync(0xffffffffu, sb[i], 8); sb[i] += __shfl_xor_sync(0xffffffffu, sb[i], 16);
  }
  if (sub == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { red[wid][c4 * 4 + i] = sw[i]; red[wid][32 + c4 * 4 + i] = sb[i]; }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    partial[(size_t)blockIdx.x * 64 + threadIdx.x] = s;
  }
}
__global__ void k_ln_reduce(const float* __restrict__ partial, int nblocks, float* __restrict__ dw, float* __restrict__ db) {
  const int c = threadIdx.x;          // 64 threads
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * 64 + c];
  if (c < 32) { if (dw) dw[c] += s; } else if (db) db[c - 32] += s;
}
constexpr int kLnBlocks = 592;

// 3-token, 3-head attention backward, thread per (point, head)      renderer.py:966-977
//   o_i = sum_j P_ij v_j, P = softmax(0.25 q k^T):  dv_j = sum_i P_ij do_i;  dP_ij = do_i . v_j;  dS_ij = P_ij (dP_ij - sum_j' P_ij' dP_ij');
//   dq_i = 0.25 sum_j dS_ij k_j;  dk_j = 0.25 sum_i dS_ij q_i
__global__ void __launch_bounds__(128) k_attention3_bwd(const float* __restrict__ qkv, const float* __restrict__ datt, float* __restrict__ dqkv, int np) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= np * 3) return;
  const int p = idx / 3, h = idx - p * 3;
  const float* base = qkv + (size_t)p * 3 * 144 + h * 16;
  float q[3][16], k[3][16], v[3][16], go[3][16];
  // every 16-float head slice is 64-byte aligned: four 16-byte loads instead of sixteen scalar ones
  auto ld16 = [](float (&dst)[16], const float* src) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src) + i);
      dst[4 * i] = t.x; dst[4 * i + 1] = t.y; dst[4 * i + 2] = t.z; dst[4 * i + 3] = t.w;
    }
  };
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    ld16(q[t], base + t * 144); ld16(k[t], base + t * 144 + 48); ld16(v[t], base + t * 144 + 96);
    ld16(go[t], datt + (size_t)(p * 3 + t) * 48 + h * 16);
  }
  float P[3][3], dS[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float s[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) a += q[i][d] * k[j][d];
      s[j] = a * 0.25f;
    }
    const float mx = fmaxf(s[0], fmaxf(s[1], s[2]));
    const float e0 = expf(s[0] - mx), e1 = expf(s[1] - mx), e2 = expf(s[2] - mx);
    const float inv = 1.f / (e0 + e1 + e2);
    P[i][0] = e0 * inv; P[i][1] = e1 * inv; P[i][2] = e2 * inv;
    float dP[3], dot = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) a += go[i][d] * v[j][d];
      dP[j] = a; dot += P[i][j] * a;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) dS[i][j] = P[i][j] * (dP[j] - dot);
  }
  float* ob = dqkv + (size_t)p * 3 * 144 + h * 16;
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      float dq[4], dk[4], dv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = 4 * d4 + e;
        dq[e] = 0.25f * (dS[t][0] * k[0][d] + dS[t][1] * k[1][d] + dS[t][2] * k[2][d]);
        dk[e] = 0.25f * (dS[0][t] * q[0][d] + dS[1][t] * q[1][d] + dS[2][t] * q[2][d]);
        dv[e] = P[0][t] * go[0][d] + P[1][t] * go[1][d] + P[2][t] * go[2][d];
      }
      reinterpret_cast<float4*>(ob + t * 144)[d4] = make_float4(dq[0], dq[1], dq[2], dq[3]);
      reinterpret_cast<float4*>(ob + t * 144 + 48)[d4] = make_float4(dk[0], dk[1], dk[2], dk[3]);
      reinterpret_cast<float4*>(ob + t * 144 + 96)[d4] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
}

// [M][C] channels-last gradient grid -> [C][M] (PyTorch layout), added into the caller's tensor is not needed: plain transpose (out = in^T)
__global__ void __launch_bounds__(256) k_from_channels_last(const float* __restrict__ in, float* __restrict__ out, int C, long long M) {
  __shared__ float tile[32][33];
  const long long m0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const long long m = m0 + r; const int c = c0 + tx;
    tile[r][tx] = (m < M && c < C) ? in[m * C + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r; const long long m = m0 + tx;
    if (c < C && m < M) out[(long long)c * M + m] = tile[tx][r];
  }
}

int run_from_channels_last(const float* in, float* out, int C, int64_t M, cudaStream_t st) {
  k_from_channels_last<<<dim3((unsigned)((M + 31) / 32), ceil_div(C, 32)), 256, 0, st>>>(in, out, C, (long long)M);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

// ------------------------------------------------------------------------------------------------- one chunk
void carve_bwd_chunk(float* base, int cap, BwdChunk& b);
// forward activations 3 156 + gradients 2 564 floats per point, + the split partial sums of the weight gradients
size_t bwd_chunk_floats(int cap) {
  BwdChunk b;
  carve_bwd_chunk(nullptr, cap, b);
  return (size_t)(b.ln_part - (float*)nullptr) + (size_t)kLnBlocks * 64 + 64;
}

void carve_bwd_chunk(float* base, int cap, BwdChunk& b) {
  const size_t c = (size_t)cap;
  float* p = base;
  b.cap = cap;
  b.comb = p; p += c * 288; b.f3raw = p; p += c * 192; b.geo = p; p += c * 8;
  b.tok = p; p += c * 96; b.ln1 = p; p += c * 96; b.qkv = p; p += c * 432; b.att = p; p += c * 144; b.tok2 = p; p += c * 96;
  b.ln2 = p; p += c * 96; b.ffp = p; p += c * 96; b.ffa = p; p += c * 96; b.tok3 = p; p += c * 96;
  b.x = p; p += c * 72; b.hb = p; p += c * 200; b.fv = p; p += c * 188; b.vh = p; p += c * 64;
  for (int i = 0; i < 8; ++i) { if (i == 4) { b.h[4] = nullptr; continue; } b.h[i] = p; p += c * 128; }
  b.dvh = p; p += c * 64; b.dpre = p; p += c * 4; b.dfv = p; p += c * 188; b.dha = p; p += c * 128; b.dhb = p; p += c * 128;
  b.dx = p; p += c * 72; b.dtok3 = p; p += c * 96; b.dff = p; p += c * 96; b.dln = p; p += c * 96; b.dtok2 = p; p += c * 96;
  b.dtok = p; p += c * 96; b.datt = p; p += c * 144; b.dqkv = p; p += c * 432; b.dcomb = p; p += c * 288; b.df3raw = p; p += c * 192;
  b.part = p; p += (size_t)(max(ceil_div(cap * 3, kRowsPerSplit) + 1, kGradWMaxSplits + 1)) * 144 * 200;
  b.ln_part = p; p += (size_t)kLnBlocks * 64;
}

// G: the forward gather parameters of this chunk (comb / f3raw / geo pointing at b.comb / b.f3raw / b.geo); rgb / dsig / drgb: per-point
// arrays of the whole pass (absolute indices).  gw entries may be NULL (no gradient wanted for that parameter).
int run_backward_chunk(const SherfWeights& w, const PackedWeights& pw, const CanonWeights& cw, const CanonBwdWeights& cbw, const SherfWeightGrads& gw,
                       GatherParams G, const BwdChunk& b, int np, int64_t p0, const float* rgb, const float* dsig, const float* drgb, cudaStream_t st) {
  if (np <= 0) return SHERF_OK;
  const int M = np, R = 3 * np;
  const int E = 256;
  const bool simt = bwd_simt();
  // one forward layer / one dX product on the selected arithmetic
  auto fwd = [&](const PackedLayer& P, const CanonLayer& C, const float* A, int lda, float* Y, int ldy, int rows, int act, const float* Res = nullptr,
                 int ldr = 0, int yg = 0, int ygs = 0) -> int {
    if (simt) return launch_simt_linear(P, A, lda, Y, ldy, rows, act, st, Res, ldr, yg, ygs);
    return launch_umma_linear(3, C, A, lda, Y, ldy, rows, act, st, Res, ldr, yg, ygs);
  };
  auto dxp = [&](const CanonLayer& C, const float* dY, int lda, const float* W, int ldb, float* dX, int ldx, int rows, int Nc, int Kr,
                 const float* Mask = nullptr, int ldm = 0, int accum = 0, int agroup = 0, int agstride = 0) -> int {
    if (simt) return gemm_nn(dY, lda, W, ldb, dX, ldx, rows, Nc, Kr, st, Mask, ldm, accum, agroup, agstride);
    return launch_umma_dx(C, dY, lda, dX, ldx, rows, st, Mask, ldm, accum, agroup, agstride);
  };
  // ---- recompute the forward with every activation kept (per-layer kernels: 3xTF32 on tcgen05, or fp32 FMA with SHERF_BWD_SIMT=1) ----
  G.comb = b.comb; G.f3raw = b.f3raw; G.geo = b.geo; G.p0 = p0; G.np = np; G.dc = DevCount{nullptr, 0, 0};
  G.dbg_vid3 = nullptr; G.dbg_can = nullptr; G.dbg_cdir = nullptr; G.dbg_uv = nullptr; G.dbg_feat = nullptr; G.dbg_max = 0; G.dbg_feat_max = 0;
  G.g_planes_cl = nullptr; G.g_feat_cl = nullptr; G.g_vol_cl[0] = G.g_vol_cl[1] = G.g_vol_cl[2] = nullptr;
  RC(run_point_gather(G, st));
  RC(fwd(pw.proj, cw.proj, b.f3raw, 192, b.comb + 64, 288, M, 0, nullptr, 0, 32, 96));
  RC(fwd(pw.reproj, cw.reproj, b.comb, 96, b.tok, 32, R, 0));
  RC(run_layernorm32(b.tok, w.ln1_w, w.ln1_b, b.ln1, R, st));
  RC(fwd(pw.qkv, cw.qkv, b.ln1, 32, b.qkv, 144, R, 0));
  RC(run_attention3(b.qkv, b.att, M, st));
  RC(fwd(pw.attn_out, cw.attn_out, b.att, 48, b.tok2, 32, R, 0, b.tok, 32));
  RC(run_layernorm32(b.tok2, w.ln2_w, w.ln2_b, b.ln2, R, st));
  RC(fwd(pw.ff1, cw.ff1, b.ln2, 32, b.ffp, 32, R, 0));
  k_gelu_fwd<<<ceil_div(R * 32, E), E, 0, st>>>(b.ffp, b.ffa, (size_t)R * 32);
  SHERF_LAUNCH_CHECK();
  RC(fwd(pw.ff2, cw.ff2, b.ffa, 32, b.tok3, 32, R, 0, b.tok2, 32));
  RC(run_decoder_inputs(b.geo, b.tok3, b.x, b.hb, b.fv, M, st));
  const float* hin[8] = {b.x, b.h[0], b.h[1], b.h[2], b.h[3], b.hb, b.h[5], b.h[6]};
  const int hld[8] = {72, 128, 128, 128, 128, 200, 128, 128};
  float* hout[8] = {b.h[0], b.h[1], b.h[2], b.h[3], b.hb + 71, b.h[5], b.h[6], b.h[7]};
  const int hold[8] = {128, 128, 128, 128, 200, 128, 128, 128};
  for (int i = 0; i < 8; ++i) RC(fwd(pw.pts[i], cw.pts[i], hin[i], hld[i], hout[i], hold[i], M, 1 /* ReLU */));
  RC(fwd(pw.feature, cw.feature, b.h[7], 128, b.fv, 188, M, 0));
  RC(fwd(pw.views, cw.views, b.fv, 188, b.vh, 64, M, 1));

  // ---- decoder backward (triplane.py:285-316) ----
  const float* dsg = dsig + p0;
  k_head_bwd<<<ceil_div(M * 64, E), E, 0, st>>>(drgb + p0 * 3, rgb + p0 * 3, b.vh, w.rgb_w, b.dpre, b.dvh, M);
  SHERF_LAUNCH_CHECK();
  RC(grad_w(b.dpre, 4, 3, b.vh, 64, 64, M, gw.rgb_w, gw.rgb_b, b.part, st));
  RC(grad_w(b.dvh, 64, 64, b.fv, 188, 187, M, gw.views_w, gw.views_b, b.part, st));
  RC(dxp(cbw.views, b.dvh, 64, w.views_w, 187, b.dfv, 188, M, 187, 64));                              // d [feature | PE4(dir) | tok1]
  RC(grad_w(b.dfv, 188, 128, b.h[7], 128, 128, M, gw.feature_w, gw.feature_b, b.part, st));
  RC(grad_w(dsg, 1, 1, b.h[7], 128, 128, M, gw.alpha_w, gw.alpha_b, b.part, st));
  RC(dxp(cbw.feature, b.dfv, 188, w.feature_w, 128, b.dha, 128, M, 128, 128));
  k_dh7<<<ceil_div(M * 128, E), E, 0, st>>>(b.dha, b.h[7], dsg, w.alpha_w, M);
  SHERF_LAUNCH_CHECK();
  float* dcur = b.dha; float* dnext = b.dhb;
  for (int i = 7; i >= 6; --i) {                                                                    // layers 7, 6: 128 -> 128
    RC(grad_w(dcur, 128, 128, hin[i], hld[i], 128, M, gw.pts_w[i], gw.pts_b[i], b.part, st));
    RC(dxp(cbw.pts[i], dcur, 128, w.pts_w[i], 128, dnext, 128, M, 128, 128, hin[i], hld[i]));
    float* t = dcur; dcur = dnext; dnext = t;
  }
  // layer 5 reads cat([x, h4]) (triplane.py:299-300): x part -> dx (no ReLU), h4 part -> dh4 masked by h4 > 0
  RC(grad_w(dcur, 128, 128, b.hb, 200, 199, M, gw.pts_w[5], gw.pts_b[5], b.part, st));
  RC(dxp(cbw.pts5x, dcur, 128, w.pts_w[5], 199, b.dx, 72, M, 71, 128));
  RC(dxp(cbw.pts[5], dcur, 128, w.pts_w[5] + 71, 199, dnext, 128, M, 128, 128, b.hb + 71, 200));
  { float* t = dcur; dcur = dnext; dnext = t; }
  for (int i = 4; i >= 1; --i) {
    RC(grad_w(dcur, 128, 128, hin[i], hld[i], 128, M, gw.pts_w[i], gw.pts_b[i], b.part, st));
    RC(dxp(cbw.pts[i], dcur, 128, w.pts_w[i], 128, dnext, 128, M, 128, 128, hin[i], hld[i]));
    float* t = dcur; dcur = dnext; dnext = t;
  }
  RC(grad_w(dcur, 128, 128, b.x, 72, 71, M, gw.pts_w[0], gw.pts_b[0], b.part, st));
  RC(dxp(cbw.pts[0], dcur, 128, w.pts_w[0], 71, b.dx, 72, M, 71, 128, nullptr, 0, 1 /* += the skip branch */));

  // ---- transformer backward (renderer.py:920-993) ----
  k_tok3_grad<<<ceil_div(M * 96, E), E, 0, st>>>(b.dx, b.dfv, b.dtok3, M);
  SHERF_LAUNCH_CHECK();
  // x3 = ff2(gelu(ff1(LN2(x2)))) + x2
  RC(grad_w(b.dtok3, 32, 32, b.ffa, 32, 32, R, gw.ff2_w, gw.ff2_b, b.part, st));
  RC(dxp(cbw.ff2, b.dtok3, 32, w.ff2_w, 32, b.dff, 32, R, 32, 32));
  k_gelu_bwd<<<ceil_div(R * 32, E), E, 0, st>>>(b.dff, b.ffp, (size_t)R * 32);
  SHERF_LAUNCH_CHECK();
  RC(grad_w(b.dff, 32, 32, b.ln2, 32, 32, R, gw.ff1_w, gw.ff1_b, b.part, st));
  RC(dxp(cbw.ff1, b.dff, 32, w.ff1_w, 32, b.dln, 32, R, 32, 32));
  k_ln_bwd<<<kLnBlocks, 256, 0, st>>>(b.tok2, w.ln2_w, b.dln, b.dtok3 /* residual branch */, b.dtok2, R, b.ln_part);
  SHERF_LAUNCH_CHECK();
  k_ln_reduce<<<1, 64, 0, st>>>(b.ln_part, kLnBlocks, gw.ln2_w, gw.ln2_b);
  SHERF_LAUNCH_CHECK();
  // x2 = to_out(attention(to_qkv(LN1(x)))) + x
  RC(grad_w(b.dtok2, 32, 32, b.att, 48, 48, R, gw.attn_out_w, gw.attn_out_b, b.part, st));
  RC(dxp(cbw.attn_out, b.dtok2, 32, w.attn_out_w, 48, b.datt, 48, R, 48, 32));
  k_attention3_bwd<<<ceil_div(M * 3, 128), 128, 0, st>>>(b.qkv, b.datt, b.dqkv, M);
  SHERF_LAUNCH_CHECK();
  RC(grad_w(b.dqkv, 144, 144, b.ln1, 32, 32, R, gw.qkv_w, nullptr, b.part, st));
  RC(dxp(cbw.qkv, b.dqkv, 144, w.qkv_w, 32, b.dln, 32, R, 32, 144));
  k_ln_bwd<<<kLnBlocks, 256, 0, st>>>(b.tok, w.ln1_w, b.dln, b.dtok2 /* residual branch */, b.dtok, R, b.ln_part);
  SHERF_LAUNCH_CHECK();
  k_ln_reduce<<<1, 64, 0, st>>>(b.ln_part, kLnBlocks, gw.ln1_w, gw.ln1_b);
  SHERF_LAUNCH_CHECK();

  // ---- fusion convolutions (renderer.py:350,423-424) ----
  RC(grad_w(b.dtok, 32, 32, b.comb, 96, 96, R, gw.reproj_w, gw.reproj_b, b.part, st));
  RC(dxp(cbw.reproj, b.dtok, 32, w.reproj_w, 96, b.dcomb, 96, R, 96, 32));
  // conv1d_projection: its 96 outputs are columns 64..95 of each token's slice
  RC(grad_w(b.dcomb + 64, 288, 96, b.f3raw, 192, 192, M, gw.proj_w, gw.proj_b, b.part, st, 32, 96));
  return SHERF_OK;
}

// second half of a chunk: input gradients (needs the gradient grids); split from run_backward_chunk so that G's grid pointers are explicit
int run_backward_chunk_inputs(const SherfWeights& w, const CanonBwdWeights& cbw, GatherParams G, const BwdChunk& b, int np, int64_t p0, cudaStream_t st) {
  if (np <= 0) return SHERF_OK;
  if (!G.g_planes_cl && !G.g_feat_cl && !G.g_vol_cl[0] && !G.g_vol_cl[1] && !G.g_vol_cl[2]) return SHERF_OK;
  if (G.g_vol_cl[0] || G.g_vol_cl[1] || G.g_vol_cl[2])
  {
    if (bwd_simt()) RC(gemm_nn(b.dcomb + 64, 288, w.proj_w, 192, b.df3raw, 192, np, 192, 96, st, nullptr, 0, 0, 32, 96));
    else RC(launch_umma_dx(cbw.proj, b.dcomb + 64, 288, b.df3raw, 192, np, st, nullptr, 0, 0, 32, 96));
  }
  G.comb = b.dcomb; G.f3raw = b.df3raw; G.geo = b.geo; G.p0 = p0; G.np = np; G.dc = DevCount{nullptr, 0, 0};
  G.dbg_vid3 = nullptr; G.dbg_can = nullptr; G.dbg_cdir = nullptr; G.dbg_uv = nullptr; G.dbg_feat = nullptr; G.dbg_max = 0; G.dbg_feat_max = 0;
  RC(run_point_scatter(G, st));
  return SHERF_OK;
}

}  // namespace sherf
