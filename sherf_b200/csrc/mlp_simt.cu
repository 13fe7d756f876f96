// Stage 3 (fp32 CUDA-core path = parity mode): feature fusion, 3-token transformer, NeRF decoder on a chunk of
// compacted points.  Replaces renderer.py:350 (conv1d_projection), :423-427 (conv1d_reprojection, Transformer
// :920-993) and triplane.py:285-316 (NeRFDecoder.forward).  fp32 FMA throughout: the reference runs with TF32
// disabled (training_loop.py:169-171).  Each linear layer is one launch of a register-tiled SGEMM whose epilogue
// fuses bias / ReLU / GELU / residual and writes straight into the next layer's (possibly concatenated) input.
#include "common.cuh"
#include "stages.cuh"

namespace sherf {

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------------- weight packing
struct PackJob { const float* w; float* wt; int N, K, kp, np; };
struct PackJobs { PackJob j[17]; int n; };

// W[N][K] (PyTorch) -> Wt[kp][np] zero padded (k-major, n contiguous)
__global__ void k_pack_weights(const PackJobs jobs) {
  const PackJob jb = jobs.j[blockIdx.y];
  const int total = jb.kp * jb.np;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / jb.np, n = i - k * jb.np;
    jb.wt[i] = (k < jb.K && n < jb.N) ? jb.w[(size_t)n * jb.K + k] : 0.f;
  }
}

size_t packed_weight_floats() {
  const int dims[17][2] = {{96, 192}, {32, 96}, {144, 32}, {32, 48}, {32, 32}, {32, 32}, {128, 71}, {128, 128}, {128, 128},
                           {128, 128}, {128, 128}, {128, 199}, {128, 128}, {128, 128}, {128, 128}, {64, 187}, {0, 0}};
  size_t t = 0;
  for (int i = 0; i < 16; ++i) t += (size_t)round_up(dims[i][1], 16) * round_up(dims[i][0], 32);
  return t;
}

int run_pack_weights(const SherfWeights& w, float* base, PackedWeights& pw, cudaStream_t st) {
  PackJobs jobs;
  jobs.n = 0;
  float* cur = base;
  auto add = [&](PackedLayer& L, const float* W, const float* b, int N, int K) {
    L.K = K; L.N = N; L.kp = round_up(K, 16); L.np = round_up(N, 32); L.wt = cur; L.bias = b;
    PackJob& j = jobs.j[jobs.n++];
    j.w = W; j.wt = cur; j.N = N; j.K = K; j.kp = L.kp; j.np = L.np;
    cur += (size_t)L.kp * L.np;
  };
  add(pw.proj, w.proj_w, w.proj_b, 96, 192);
  add(pw.reproj, w.reproj_w, w.reproj_b, 32, 96);
  add(pw.qkv, w.qkv_w, nullptr, 144, 32);
  add(pw.attn_out, w.attn_out_w, w.attn_out_b, 32, 48);
  add(pw.ff1, w.ff1_w, w.ff1_b, 32, 32);
  add(pw.ff2, w.ff2_w, w.ff2_b, 32, 32);
  const int ptsK[8] = {71, 128, 128, 128, 128, 199, 128, 128};
  for (int i = 0; i < 8; ++i) add(pw.pts[i], w.pts_w[i], w.pts_b[i], 128, ptsK[i]);
  add(pw.feature, w.feature_w, w.feature_b, 128, 128);
  add(pw.views, w.views_w, w.views_b, 64, 187);
  if (!g_pack_plan_only) {
    k_pack_weights<<<dim3(8, jobs.n), 256, 0, st>>>(jobs);
    SHERF_LAUNCH_CHECK();
  }
  return SHERF_OK;
}

// ------------------------------------------------------------------------------------------------- SGEMM
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

struct GemmArgs {
  const float* A; int lda;            // [M][lda] row-major, columns [0,K)
  const float* Wt; int np;            // packed [kp][np]
  const float* bias;                  // [N] or null
  float* Y; int ldy;                  // output rows
  int ygroup, ygstride;               // column c -> (c / ygroup) * ygstride + c % ygroup   (ygroup = 0: identity)
  const float* Res; int ldr;          // optional residual, added after the activation
  int act;
  int M, N, K;
};

// BM=128 rows x BN cols per CTA, BK=16, 256 threads, thread tile 8 x (BN/16).
template <int BN>
__global__ void __launch_bounds__(256) k_sgemm(const GemmArgs g) {
  constexpr int BM = 128, BK = 16, TM = 8, TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // A loader: thread -> row (tid & 127), k-group (tid >> 7) and +2   (two float4 per thread per tile)
  const int arow = tid & 127, akg = tid >> 7;
  const bool arow_ok = (m0 + arow) < g.M;
  const float* aptr = g.A + (size_t)(m0 + arow) * g.lda;
  // B loader: BK x BN floats = 16*BN/4 float4; thread loads (16*BN/4)/256 float4
  constexpr int BV = (BK * BN / 4) / 256 > 0 ? (BK * BN / 4) / 256 : 1;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2];
  float4 rb[BV];
  const int ktiles = (g.K + BK - 1) / BK;
  auto load_tile = [&](int kt) {
    const int k0 = kt * BK;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = k0 + (akg + 2 * h) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (arow_ok && k < g.K) {
        if (k + 3 < g.K) v = *reinterpret_cast<const float4*>(aptr + k);   // lda % 4 == 0 and base 16 B aligned
        else { v.x = aptr[k]; if (k + 1 < g.K) v.y = aptr[k + 1]; if (k + 2 < g.K) v.z = aptr[k + 2]; }
      }
      ra[h] = v;
    }
#pragma unroll
    for (int h = 0; h < BV; ++h) {
      const int idx = tid + h * 256;                 // float4 index within the tile
      if (idx < BK * BN / 4) {
        const int k = idx / (BN / 4), c4 = idx % (BN / 4);
        rb[h] = *reinterpret_cast<const float4*>(g.Wt + (size_t)(k0 + k) * g.np + n0 + c4 * 4);   // kp, np padded with zeros
      }
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = (akg + 2 * h) * 4;
      As[kk + 0][arow] = ra[h].x; As[kk + 1][arow] = ra[h].y; As[kk + 2][arow] = ra[h].z; As[kk + 3][arow] = ra[h].w;
    }
#pragma unroll
    for (int h = 0; h < BV; ++h) {
      const int idx = tid + h * 256;
      if (idx < BK * BN / 4) {
        const int k = idx / (BN / 4), c4 = idx % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[k][c4 * 4]) = rb[h];
      }
    }
  };
  load_tile(0);
  for (int kt = 0; kt < ktiles; ++kt) {
    __syncthreads();
    store_tile();
    __syncthreads();
    if (kt + 1 < ktiles) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j] + (g.bias ? g.bias[n] : 0.f);
      if (g.act == ACT_RELU) v = fmaxf(v, 0.f);
      else if (g.act == ACT_GELU) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
      if (g.Res) v += g.Res[(size_t)m * g.ldr + n];
      const int col = g.ygroup ? (n / g.ygroup) * g.ygstride + (n % g.ygroup) : n;
      g.Y[(size_t)m * g.ldy + col] = v;
    }
  }
}

int launch_simt_linear(const PackedLayer& L, const float* A, int lda, float* Y, int ldy, int M, int act, cudaStream_t st,
                       const float* Res, int ldr, int ygroup, int ygstride) {
  GemmArgs g;
  g.A = A; g.lda = lda; g.Wt = L.wt; g.np = L.np; g.bias = L.bias; g.Y = Y; g.ldy = ldy; g.ygroup = ygroup; g.ygstride = ygstride;
  g.Res = Res; g.ldr = ldr; g.act = act; g.M = M; g.N = L.N; g.K = L.K;
  if (L.np % 128 == 0) {
    k_sgemm<128><<<dim3(ceil_div(M, 128), L.np / 128), 256, 0, st>>>(g);
  } else if (L.np % 64 == 0) {
    k_sgemm<64><<<dim3(ceil_div(M, 128), L.np / 64), 256, 0, st>>>(g);
  } else {
    k_sgemm<32><<<dim3(ceil_div(M, 128), L.np / 32), 256, 0, st>>>(g);
  }
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

// ------------------------------------------------------------------------------------------------- small kernels
// LayerNorm over 32 channels, eps 1e-5, warp per row (renderer.py:931)
__global__ void k_layernorm32(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                              float* __restrict__ y, int rows) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float v = x[(size_t)r * 32 + lane];
  float s = v;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / 32.f);
  const float d = v - mean;
  float q = d * d;
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / 32.f) + 1e-5f);
  y[(size_t)r * 32 + lane] = d * rstd * w[lane] + b[lane];
}

// 3-token, 3-head (dim 16) attention; thread per (point, head).  qkv rows = p*3+tok, [q(48)|k(48)|v(48)]   renderer.py:966-977
__global__ void __launch_bounds__(128) k_attention3(const float* __restrict__ qkv, float* __restrict__ att, int np) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= np * 3) return;
  const int p = idx / 3, h = idx - p * 3;
  const float* base = qkv + (size_t)p * 3 * 144 + h * 16;
  float dots[3][3];
  {
    float q[3][16], k[3][16];
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        const float4 a = *reinterpret_cast<const float4*>(base + t * 144 + d4 * 4);
        const float4 b = *reinterpret_cast<const float4*>(base + t * 144 + 48 + d4 * 4);
        q[t][d4 * 4] = a.x; q[t][d4 * 4 + 1] = a.y; q[t][d4 * 4 + 2] = a.z; q[t][d4 * 4 + 3] = a.w;
        k[t][d4 * 4] = b.x; k[t][d4 * 4 + 1] = b.y; k[t][d4 * 4 + 2] = b.z; k[t][d4 * 4 + 3] = b.w;
      }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) s += q[i][d] * k[j][d];
        dots[i][j] = s * 0.25f;                             // dim_head ** -0.5
      }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float mx = fmaxf(dots[i][0], fmaxf(dots[i][1], dots[i][2]));
    const float e0 = expf(dots[i][0] - mx), e1 = expf(dots[i][1] - mx), e2 = expf(dots[i][2] - mx);
    const float inv = 1.f / (e0 + e1 + e2);
    dots[i][0] = e0 * inv; dots[i][1] = e1 * inv; dots[i][2] = e2 * inv;
  }
#pragma unroll
  for (int d4 = 0; d4 < 4; ++d4) {
    const float4 v0 = *reinterpret_cast<const float4*>(base + 0 * 144 + 96 + d4 * 4);
    const float4 v1 = *reinterpret_cast<const float4*>(base + 1 * 144 + 96 + d4 * 4);
    const float4 v2 = *reinterpret_cast<const float4*>(base + 2 * 144 + 96 + d4 * 4);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float4 o;
      o.x = dots[i][0] * v0.x + dots[i][1] * v1.x + dots[i][2] * v2.x;
      o.y = dots[i][0] * v0.y + dots[i][1] * v1.y + dots[i][2] * v2.y;
      o.z = dots[i][0] * v0.z + dots[i][1] * v1.z + dots[i][2] * v2.z;
      o.w = dots[i][0] * v0.w + dots[i][1] * v1.w + dots[i][2] * v2.w;
      *reinterpret_cast<float4*>(att + (size_t)(p * 3 + i) * 48 + h * 16 + d4 * 4) = o;
    }
  }
}

// Decoder inputs: x = [PE6(can) | tok0] (71) into x[72] and hb[0:71]; fv[128:187] = [PE4(cdir) | tok1].   renderer.py:432, triplane.py:293-310
__global__ void k_decoder_inputs(const float* __restrict__ geo, const float* __restrict__ tok3, float* __restrict__ x,
                                 float* __restrict__ hb, float* __restrict__ fv, int np, float* dbg_tok, int64_t p0, int64_t dbg_max) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= np) return;
  const float g = lane < 8 ? geo[(size_t)p * 8 + lane] : 0.f;
  // broadcast the six geometry values with full-warp shuffles BEFORE any divergent code
  float gv[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) gv[k] = __shfl_sync(0xffffffffu, g, k);
  // positional encodings: out[0:3] = v ; out[3 + m*3 + c] = sin(phase(m) + v[c] * 2^(m/2))       renderer.py:900-916
  auto pe = [&](int e, float v0, float v1, float v2) -> float {
    if (e < 3) return e == 0 ? v0 : (e == 1 ? v1 : v2);
    const int m = (e - 3) / 3, c = (e - 3) - 3 * m;
    const float vv = c == 0 ? v0 : (c == 1 ? v1 : v2);
    return sinf(__fadd_rn((m & 1) ? kPi2 : 0.f, __fmul_rn(vv, (float)(1 << (m >> 1)))));
  };
  for (int e = lane; e < 39; e += 32) {
    const float val = pe(e, gv[0], gv[1], gv[2]);
    x[(size_t)p * 72 + e] = val;
    hb[(size_t)p * 200 + e] = val;
  }
  if (lane < 27) fv[(size_t)p * 188 + 128 + lane] = pe(lane, gv[3], gv[4], gv[5]);
  const float t0 = tok3[(size_t)(p * 3 + 0) * 32 + lane], t1 = tok3[(size_t)(p * 3 + 1) * 32 + lane];
  x[(size_t)p * 72 + 39 + lane] = t0;
  hb[(size_t)p * 200 + 39 + lane] = t0;
  fv[(size_t)p * 188 + 155 + lane] = t1;
  if (lane == 0) { x[(size_t)p * 72 + 71] = 0.f; fv[(size_t)p * 188 + 187] = 0.f; hb[(size_t)p * 200 + 199] = 0.f; }
  if (dbg_tok && p0 + p < dbg_max) { dbg_tok[(p0 + p) * 64 + lane] = t0; dbg_tok[(p0 + p) * 64 + 32 + lane] = t1; }
}

// Transformer tail in one pass, thread per (point, query token t in {0,1}) -- token 2 is only ever a key/value source and
// the decoder never reads its output (triplane.py:288-289).  Per thread: 3-head attention of query t over the 3 tokens,
// to_out projection + residual, LayerNorm, FeedForward (Linear-GELU-Linear) + residual (renderer.py:966-993), then the
// decoder inputs: t = 0 -> x = [PE6(can) | tok0], t = 1 -> [PE4(cdir) | tok1] (renderer.py:432, triplane.py:293,308).
// All fp32 FMA; weights are broadcast from shared memory.
struct TailParams {
  const float *qkv, *tok, *geo;                      // [3np][144], [3np][32] (pre-attention tokens), [np][8]
  const float *wo, *bo, *ln_w, *ln_b, *w1, *b1, *w2, *b2;
  float *x, *hb, *fv;                                // [np][72], optional [np][200], [np][188]
  float* dbg_tok; int64_t p0, dbg_max;
  int np;
};

__global__ void __launch_bounds__(128) k_transformer_tail(const TailParams P) {
  __shared__ __align__(16) float s_wo[32 * 48];
  __shared__ __align__(16) float s_w1[32 * 32];
  __shared__ __align__(16) float s_w2[32 * 32];
  __shared__ float s_bo[32], s_lnw[32], s_lnb[32], s_b1[32], s_b2[32];
  for (int i = threadIdx.x; i < 32 * 48; i += blockDim.x) s_wo[i] = P.wo[i];
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) { s_w1[i] = P.w1[i]; s_w2[i] = P.w2[i]; }
  if (threadIdx.x < 32) {
    s_bo[threadIdx.x] = P.bo[threadIdx.x]; s_lnw[threadIdx.x] = P.ln_w[threadIdx.x]; s_lnb[threadIdx.x] = P.ln_b[threadIdx.x];
    s_b1[threadIdx.x] = P.b1[threadIdx.x]; s_b2[threadIdx.x] = P.b2[threadIdx.x];
  }
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  // warp-uniform query token: warps alternate t = 0 / 1 over the same 32 points (no divergence, K/V rows shared through L1)
  const int p = (idx >> 6) * 32 + (idx & 31), t = (idx >> 5) & 1;
  if (p >= P.np) return;
  const float* qrow = P.qkv + (size_t)(p * 3 + t) * 144;
  float out[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) out[o] = s_bo[o];
#pragma unroll 1
  for (int h = 0; h < 3; ++h) {
    float q[16], att[16];
    float dots[3];
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      const float4 a = *reinterpret_cast<const float4*>(qrow + h * 16 + d4 * 4);
      q[d4 * 4] = a.x; q[d4 * 4 + 1] = a.y; q[d4 * 4 + 2] = a.z; q[d4 * 4 + 3] = a.w;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float* krow = P.qkv + (size_t)(p * 3 + j) * 144 + 48 + h * 16;
      float sdot = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        const float4 b = *reinterpret_cast<const float4*>(krow + d4 * 4);
        sdot += q[d4 * 4] * b.x; sdot += q[d4 * 4 + 1] * b.y; sdot += q[d4 * 4 + 2] * b.z; sdot += q[d4 * 4 + 3] * b.w;
      }
      dots[j] = sdot * 0.25f;                               // dim_head ** -0.5
    }
    const float mx = fmaxf(dots[0], fmaxf(dots[1], dots[2]));
    const float e0 = expf(dots[0] - mx), e1 = expf(dots[1] - mx), e2 = expf(dots[2] - mx);
    const float inv = 1.f / (e0 + e1 + e2);
    const float a0 = e0 * inv, a1 = e1 * inv, a2 = e2 * inv;
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      const float4 v0 = *reinterpret_cast<const float4*>(P.qkv + (size_t)(p * 3 + 0) * 144 + 96 + h * 16 + d4 * 4);
      const float4 v1 = *reinterpret_cast<const float4*>(P.qkv + (size_t)(p * 3 + 1) * 144 + 96 + h * 16 + d4 * 4);
      const float4 v2 = *reinterpret_cast<const float4*>(P.qkv + (size_t)(p * 3 + 2) * 144 + 96 + h * 16 + d4 * 4);
      att[d4 * 4 + 0] = a0 * v0.x + a1 * v1.x + a2 * v2.x;
      att[d4 * 4 + 1] = a0 * v0.y + a1 * v1.y + a2 * v2.y;
      att[d4 * 4 + 2] = a0 * v0.z + a1 * v1.z + a2 * v2.z;
      att[d4 * 4 + 3] = a0 * v0.w + a1 * v1.w + a2 * v2.w;
    }
    // to_out: out[o] += sum_d att[d] * Wo[o][h*16 + d]
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      const float4* wr = reinterpret_cast<const float4*>(s_wo + o * 48 + h * 16);
      float acc = out[o];
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        const float4 wv = wr[d4];
        acc = fmaf(att[d4 * 4], wv.x, acc); acc = fmaf(att[d4 * 4 + 1], wv.y, acc);
        acc = fmaf(att[d4 * 4 + 2], wv.z, acc); acc = fmaf(att[d4 * 4 + 3], wv.w, acc);
      }
      out[o] = acc;
    }
  }
  // residual, LayerNorm (eps 1e-5)
  float tok2[32], ln[32];
  {
    const float4* tr = reinterpret_cast<const float4*>(P.tok + (size_t)(p * 3 + t) * 32);
    float mean = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < 8; ++d4) {
      const float4 tv = tr[d4];
      tok2[d4 * 4] = out[d4 * 4] + tv.x; tok2[d4 * 4 + 1] = out[d4 * 4 + 1] + tv.y;
      tok2[d4 * 4 + 2] = out[d4 * 4 + 2] + tv.z; tok2[d4 * 4 + 3] = out[d4 * 4 + 3] + tv.w;
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) mean += tok2[o];
    mean *= (1.f / 32.f);
    float var = 0.f;
#pragma unroll
    for (int o = 0; o < 32; ++o) { const float d = tok2[o] - mean; var += d * d; }
    const float rstd = rsqrtf(var * (1.f / 32.f) + 1e-5f);
#pragma unroll
    for (int o = 0; o < 32; ++o) ln[o] = (tok2[o] - mean) * rstd * s_lnw[o] + s_lnb[o];
  }
  // FeedForward: Linear(32,32) - GELU(erf) - Linear(32,32), + residual
  float hmid[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    const float4* wr = reinterpret_cast<const float4*>(s_w1 + o * 32);
    float acc = s_b1[o];
#pragma unroll
    for (int d4 = 0; d4 < 8; ++d4) {
      const float4 wv = wr[d4];
      acc = fmaf(ln[d4 * 4], wv.x, acc); acc = fmaf(ln[d4 * 4 + 1], wv.y, acc);
      acc = fmaf(ln[d4 * 4 + 2], wv.z, acc); acc = fmaf(ln[d4 * 4 + 3], wv.w, acc);
    }
    hmid[o] = 0.5f * acc * (1.f + erff(acc * 0.70710678118654752440f));
  }
  float tok3[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    const float4* wr = reinterpret_cast<const float4*>(s_w2 + o * 32);
    float acc = s_b2[o];
#pragma unroll
    for (int d4 = 0; d4 < 8; ++d4) {
      const float4 wv = wr[d4];
      acc = fmaf(hmid[d4 * 4], wv.x, acc); acc = fmaf(hmid[d4 * 4 + 1], wv.y, acc);
      acc = fmaf(hmid[d4 * 4 + 2], wv.z, acc); acc = fmaf(hmid[d4 * 4 + 3], wv.w, acc);
    }
    tok3[o] = acc + tok2[o];
  }
  // decoder inputs, assembled in registers and written with 16-byte stores (rows are 16-byte aligned)
  const float g0 = P.geo[(size_t)p * 8 + 3 * t], g1 = P.geo[(size_t)p * 8 + 3 * t + 1], g2 = P.geo[(size_t)p * 8 + 3 * t + 2];
  auto pe = [&](int m, float gv) -> float {
    return sinf(__fadd_rn((m & 1) ? kPi2 : 0.f, __fmul_rn(gv, (float)(1 << (m >> 1)))));
  };
  if (t == 0) {
    float vals[72];
    vals[0] = g0; vals[1] = g1; vals[2] = g2;
#pragma unroll
    for (int m = 0; m < 12; ++m) { vals[3 + 3 * m] = pe(m, g0); vals[4 + 3 * m] = pe(m, g1); vals[5 + 3 * m] = pe(m, g2); }
#pragma unroll
    for (int o = 0; o < 32; ++o) vals[39 + o] = tok3[o];
    vals[71] = 0.f;
    float4* dx = reinterpret_cast<float4*>(P.x + (size_t)p * 72);
#pragma unroll
    for (int i = 0; i < 18; ++i) dx[i] = make_float4(vals[4 * i], vals[4 * i + 1], vals[4 * i + 2], vals[4 * i + 3]);
    if (P.hb) {
      float4* dh = reinterpret_cast<float4*>(P.hb + (size_t)p * 200);
#pragma unroll
      for (int i = 0; i < 17; ++i) dh[i] = make_float4(vals[4 * i], vals[4 * i + 1], vals[4 * i + 2], vals[4 * i + 3]);
      float* dhs = P.hb + (size_t)p * 200;
      dhs[68] = vals[68]; dhs[69] = vals[69]; dhs[70] = vals[70];      // hb[71..198] belongs to pts_linears[4]'s output
      dhs[199] = 0.f;
    }
  } else {
    float vals[60];
    vals[0] = g0; vals[1] = g1; vals[2] = g2;
#pragma unroll
    for (int m = 0; m < 8; ++m) { vals[3 + 3 * m] = pe(m, g0); vals[4 + 3 * m] = pe(m, g1); vals[5 + 3 * m] = pe(m, g2); }
#pragma unroll
    for (int o = 0; o < 32; ++o) vals[27 + o] = tok3[o];
    vals[59] = 0.f;
    float4* dv = reinterpret_cast<float4*>(P.fv + (size_t)p * 188 + 128);
#pragma unroll
    for (int i = 0; i < 15; ++i) dv[i] = make_float4(vals[4 * i], vals[4 * i + 1], vals[4 * i + 2], vals[4 * i + 3]);
  }
  if (P.dbg_tok && P.p0 + p < P.dbg_max) {
#pragma unroll
    for (int o = 0; o < 32; ++o) P.dbg_tok[(P.p0 + p) * 64 + t * 32 + o] = tok3[o];
  }
}

// sigma[p] = h[p] . w + b  (alpha_linear, triplane.py:302); warp per row of 128
__global__ void k_alpha(const float* __restrict__ h, int ldh, const float* __restrict__ w, const float* __restrict__ b,
                        float* __restrict__ sigma, int np) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= np) return;
  const float4 a = *reinterpret_cast<const float4*>(h + (size_t)p * ldh + lane * 4);
  float s = a.x * w[lane * 4] + a.y * w[lane * 4 + 1] + a.z * w[lane * 4 + 2] + a.w * w[lane * 4 + 3];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) sigma[p] = s + b[0];
}

// rgb[p] = sigmoid(vh[p] . W^T + b) * 1.002 - 0.001   (rgb_linear + sigmoid clamp, triplane.py:313-314)
__global__ void k_rgb_head(const float* __restrict__ vh, const float* __restrict__ w, const float* __restrict__ b,
                           float* __restrict__ rgb, int np) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= np) return;
  const float a0 = vh[(size_t)p * 64 + lane], a1 = vh[(size_t)p * 64 + 32 + lane];
  float s[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    s[c] = a0 * w[c * 64 + lane] + a1 * w[c * 64 + 32 + lane];
    for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
  }
  if (lane < 3) {
    const float z = (lane == 0 ? s[0] : (lane == 1 ? s[1] : s[2])) + b[lane];
    rgb[(size_t)p * 3 + lane] = (1.f / (1.f + expf(-z))) * (1.f + 2.f * 0.001f) - 0.001f;
  }
}

// ------------------------------------------------------------------------------------------------- chunk buffers
size_t chunk_buffer_floats(int cap) {
  const size_t c = (size_t)cap;
  return c * (288 + 192 + 8 + 72 + 128 + 128 + 200 + 188 + 64) + 3 * c * (32 + 32 + 144 + 48 + 32 + 32 + 32);
}
void carve_chunk_buffers(float* base, int cap, ChunkBuffers& cb) {
  const size_t c = (size_t)cap;
  float* p = base;
  cb.cap = cap;
  cb.comb = p; p += c * 288;
  cb.f3raw = p; p += c * 192;
  cb.geo = p; p += c * 8;
  cb.tok = p; p += 3 * c * 32;
  cb.ln = p; p += 3 * c * 32;
  cb.qkv = p; p += 3 * c * 144;
  cb.att = p; p += 3 * c * 48;
  cb.tok2 = p; p += 3 * c * 32;
  cb.ffh = p; p += 3 * c * 32;
  cb.tok3 = p; p += 3 * c * 32;
  cb.x = p; p += c * 72;
  cb.h1 = p; p += c * 128;
  cb.h2 = p; p += c * 128;
  cb.hb = p; p += c * 200;
  cb.fv = p; p += c * 188;
  cb.vh = p; p += c * 64;
}

// host launchers of the small fp32 kernels for the recompute pass of the backward (backward.cu)
int run_layernorm32(const float* x, const float* w, const float* b, float* y, int rows, cudaStream_t st) {
  k_layernorm32<<<ceil_div(rows, 8), 256, 0, st>>>(x, w, b, y, rows);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}
int run_attention3(const float* qkv, float* att, int np, cudaStream_t st) {
  k_attention3<<<ceil_div(np * 3, 128), 128, 0, st>>>(qkv, att, np);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}
int run_decoder_inputs(const float* geo, const float* tok3, float* x, float* hb, float* fv, int np, cudaStream_t st) {
  k_decoder_inputs<<<ceil_div(np, 8), 256, 0, st>>>(geo, tok3, x, hb, fv, np, nullptr, 0, 0);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

#define RC(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

int run_mlp(int prec, const SherfWeights& w, const PackedWeights& pw, const CanonWeights& cw, const FusedPlan* fused, const ChunkBuffers& cb, int np,
            int64_t p0, float* sigma_out, float* rgb_out, float* dbg_tok, int64_t dbg_max, cudaStream_t st, void (*span_begin)(int),
            void (*span_end)(), DevCount dc) {
  if (np <= 0) return SHERF_OK;
  const int rows3 = 3 * np;
  // SHERF_MLP_BF16X3: the decoder runs as bf16 split products (decoder_pp.cu); fusion conv and transformer stay 3xTF32
  const bool pp = prec == SHERF_MLP_BF16X3 && fused && fused->pp;
  if (prec == SHERF_MLP_BF16X3) prec = SHERF_MLP_TF32X3;
  // one linear layer on the selected arithmetic
  auto launch_gemm = [&](const PackedLayer& P, const CanonLayer& C, const float* A, int lda, float* Y, int ldy, int M, int act,
                         cudaStream_t s, const float* Res = nullptr, int ldr = 0, int yg = 0, int ygs = 0) -> int {
    if (prec == SHERF_MLP_FP32) return launch_simt_linear(P, A, lda, Y, ldy, M, act, s, Res, ldr, yg, ygs);
    return launch_umma_linear(prec == SHERF_MLP_TF32X3 ? 3 : 1, C, A, lda, Y, ldy, M, act, s, Res, ldr, yg, ygs);
  };
  if (pp && fused->fr_blob && fused->xb_blob) {
    // the front kernel (front_fused.cu) already left the tokens in cb.tok; LayerNorm-1 happens in the transformer kernel
  } else if (fused && prec != SHERF_MLP_FP32 && fused->ff_blob) {
    // conv1d_projection + conv1d_reprojection + LayerNorm-1 in one persistent tcgen05 kernel (fusion_fused.cu)
    if (span_begin) span_begin(7);
    RC(run_fusion_fused(prec == SHERF_MLP_TF32X3 ? 3 : 1, w, fused->ff_blob, cb.f3raw, cb.comb, cb.tok, cb.ln, np, st));
    if (span_end) span_end();
  } else {
    // conv1d_projection 192 -> 96, written as the third 32-wide slice of each token's 96-wide fusion input (renderer.py:350,423)
    RC(launch_gemm(pw.proj, cw.proj, cb.f3raw, 192, cb.comb + 64, 288, np, ACT_NONE, st, nullptr, 0, 32, 96));
    // conv1d_reprojection 96 -> 32 per token (renderer.py:424): rows = (point, token)
    if (prec == SHERF_MLP_FP32) {
      RC(launch_gemm(pw.reproj, cw.reproj, cb.comb, 96, cb.tok, 32, rows3, ACT_NONE, st));
      // transformer layer (renderer.py:980-993): x = attn(LN(x)) + x ; x = ff(LN(x)) + x
      k_layernorm32<<<ceil_div(rows3, 8), 256, 0, st>>>(cb.tok, w.ln1_w, w.ln1_b, cb.ln, rows3);
      SHERF_LAUNCH_CHECK();
    } else {
      // tensor-core path: the first LayerNorm is fused into the reprojection's copy-out
      RC(launch_umma_linear(prec == SHERF_MLP_TF32X3 ? 3 : 1, cw.reproj, cb.comb, 96, cb.tok, 32, rows3, ACT_NONE, st, nullptr, 0, 0, 0,
                            w.ln1_w, w.ln1_b, cb.ln, 32));
    }
  }
  if (pp && fused->xb_blob) {
    // bf16x3: LayerNorm-1 + qkv -> attention -> to_out -> LN2 -> FeedForward -> packed decoder inputs, two CTAs per SM (xformer_bf16.cu)
    if (span_begin) span_begin(6);
    RC(run_xformer_bf16(w, fused->xb_blob, cb.tok, cb.geo, np, dbg_tok, p0, dbg_max, st, fused->pp->xp, fused->pp->vp, cb.qkv /* PE scratch */, dc));
    if (span_end) span_end();
  } else if (fused && prec != SHERF_MLP_FP32 && fused->xf_blob) {
    // qkv -> attention -> to_out -> LN2 -> FeedForward -> decoder inputs in one persistent tcgen05 kernel (xformer_fused.cu)
    if (span_begin) span_begin(6);
    RC(run_xformer_fused(prec == SHERF_MLP_TF32X3 ? 3 : 1, w, fused->xf_blob, cb.ln, cb.tok, cb.geo, cb.x, cb.fv, np, dbg_tok, p0, dbg_max, st,
                         pp ? fused->pp->xp : nullptr, pp ? fused->pp->vp : nullptr, cb.qkv /* unused by the fused path: PE scratch */));
    if (span_end) span_end();
  } else {
    RC(launch_gemm(pw.qkv, cw.qkv, cb.ln, 32, cb.qkv, 144, rows3, ACT_NONE, st));
    {
      TailParams T;
      T.qkv = cb.qkv; T.tok = cb.tok; T.geo = cb.geo;
      T.wo = w.attn_out_w; T.bo = w.attn_out_b; T.ln_w = w.ln2_w; T.ln_b = w.ln2_b; T.w1 = w.ff1_w; T.b1 = w.ff1_b; T.w2 = w.ff2_w; T.b2 = w.ff2_b;
      T.x = cb.x; T.hb = (fused && prec != SHERF_MLP_FP32) ? nullptr : cb.hb; T.fv = cb.fv;
      T.dbg_tok = dbg_tok; T.p0 = p0; T.dbg_max = dbg_max; T.np = np;
      k_transformer_tail<<<ceil_div(np, 64), 128, 0, st>>>(T);       // 128 threads = 64 points x 2 query tokens
      SHERF_LAUNCH_CHECK();
    }
  }
  if (pp) {
    // whole NeRFDecoder, two tiles in flight per SM, activations in tensor memory
    if (span_begin) span_begin(5);
    if (!fused->xf_blob && !fused->xb_blob) RC(run_pack_xv(cb.x, 72, cb.fv, 188, np, fused->pp->xp, fused->pp->vp, st));   // else the transformer kernel wrote the packed tiles
    RC(run_decoder_pp(*fused->pp, w, fused->pp->xp, fused->pp->vp, sigma_out + p0, rgb_out + p0 * 3, np, st, dc));
    if (span_end) span_end();
    return SHERF_OK;
  }
  if (fused && prec != SHERF_MLP_FP32) {
    // pts_linears[0..7] + feature_linear + alpha_linear in one persistent tcgen05 kernel, activations on-chip
    if (span_begin) span_begin(5);
    RC(run_decoder_fused_plan(prec == SHERF_MLP_TF32X3 ? 3 : 1, *fused, cb.x, 72, cb.fv, 188, sigma_out + p0, rgb_out + p0 * 3, w.rgb_w, w.rgb_b, np, st));
    if (span_end) span_end();
    return SHERF_OK;                 // views_linear + rgb head are part of the fused kernel
  } else {
    RC(launch_gemm(pw.pts[0], cw.pts[0], cb.x, 72, cb.h1, 128, np, ACT_RELU, st));
    RC(launch_gemm(pw.pts[1], cw.pts[1], cb.h1, 128, cb.h2, 128, np, ACT_RELU, st));
    RC(launch_gemm(pw.pts[2], cw.pts[2], cb.h2, 128, cb.h1, 128, np, ACT_RELU, st));
    RC(launch_gemm(pw.pts[3], cw.pts[3], cb.h1, 128, cb.h2, 128, np, ACT_RELU, st));
    RC(launch_gemm(pw.pts[4], cw.pts[4], cb.h2, 128, cb.hb + 71, 200, np, ACT_RELU, st));       // skip: h = cat([x, h])  (i == 4)
    RC(launch_gemm(pw.pts[5], cw.pts[5], cb.hb, 200, cb.h1, 128, np, ACT_RELU, st));
    RC(launch_gemm(pw.pts[6], cw.pts[6], cb.h1, 128, cb.h2, 128, np, ACT_RELU, st));
    RC(launch_gemm(pw.pts[7], cw.pts[7], cb.h2, 128, cb.h1, 128, np, ACT_RELU, st));
    k_alpha<<<ceil_div(np, 8), 256, 0, st>>>(cb.h1, 128, w.alpha_w, w.alpha_b, sigma_out + p0, np);
    SHERF_LAUNCH_CHECK();
    RC(launch_gemm(pw.feature, cw.feature, cb.h1, 128, cb.fv, 188, np, ACT_NONE, st));
  }
  RC(launch_gemm(pw.views, cw.views, cb.fv, 188, cb.vh, 64, np, ACT_RELU, st));
  k_rgb_head<<<ceil_div(np, 8), 256, 0, st>>>(cb.vh, w.rgb_w, w.rgb_b, rgb_out + p0 * 3, np);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

}  // namespace sherf
