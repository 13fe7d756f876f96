// Tensor-core linear layer for the fusion / transformer / decoder stack: Y = act(A * W^T + b) (+ residual) with
// tcgen05.mma kind::tf32 (fp32 accumulate in TMEM).  Two arithmetic modes:
//   TF32   : one MMA per k-step on round-to-nearest TF32 operands                       (fast, ~1e-3 relative)
//   TF32X3 : error-compensated split  a = a_hi + a_lo, w = w_hi + w_lo;  a_lo*w_hi + a_hi*w_lo + a_hi*w_hi
//            = fp32-grade products on the tensor cores (the dropped a_lo*w_lo term is 2^-22 relative)
// Operands sit in shared memory in the K-major no-swizzle canonical layout (8-row x 16-byte core matrices):
//   A tile  [128 rows x 64 k]  : element (r,k) at (k/4)*2048 + r*16 + (k%4)*4 bytes   (LBO 2048, SBO 128)
//   W chunk [Np rows  x 64 k]  : element (n,k) at (k/4)*Np*16 + n*16 + (k%4)*4 bytes  (LBO Np*16, SBO 128)
// Weights are pre-packed in exactly that order in global memory, so a chunk is one contiguous copy.
#include "common.cuh"
#include "stages.cuh"
#include "umma.cuh"

namespace sherf {

constexpr int kKC = 64;                 // k per shared-memory chunk (16 core-matrix columns)

// W[N][K] (PyTorch) -> canonical chunks, hi (tf32-rounded) and lo (tf32-rounded residual) parts.
// trans = 0: W is [N][K] with row stride ldw (PyTorch [out][in]); trans = 1: the packed operand is W^T, i.e. element (n, k) = w[k * ldw + n]
// (the backward's dX = dY . W needs the weight matrix with the roles of its two axes swapped).  ldw = 0 means K (dense [N][K]).
struct CanonJob { const float* w; float* hi; float* lo; int N, K, Np, nchunks, ldw, trans; };
struct CanonJobs { CanonJob j[18]; int n; };

__global__ void k_pack_canonical(const CanonJobs jobs) {
  const CanonJob jb = jobs.j[blockIdx.y];
  const int total = jb.nchunks * 16 * jb.Np * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 3, n = (i >> 2) % jb.Np, kg = (i >> 2) / jb.Np;     // kg counts core-matrix columns over all chunks
    const int k = kg * 4 + e;
    const int ldw = jb.ldw ? jb.ldw : jb.K;
    float v = 0.f;
    if (n < jb.N && k < jb.K) {
      if (jb.trans >= 2) {
        // sparse-convolution input gradient: n = input channel, k = (kernel offset o, output channel co); the source is spconv's KRSC
        // weight [c_out][27][c_in]; trans == 2 reads the mirrored offset 26 - o (SubMConv3d: the output that saw input p through offset o
        // sits at p - (o - 1), i.e. at neighbour slot 26 - o of p)
        const int cout = jb.K / 27, o = k / cout, co = k - o * cout;
        v = jb.w[((size_t)co * 27 + (jb.trans == 2 ? 26 - o : o)) * jb.N + n];
      } else v = jb.trans ? jb.w[(size_t)k * ldw + n] : jb.w[(size_t)n * ldw + k];
    }
    const float h = umma::to_tf32(v);
    jb.hi[i] = h;
    jb.lo[i] = umma::to_tf32(v - h);
  }
}

struct UmmaArgs {
  const float* A; int lda;
  const float* Whi; const float* Wlo; int Np; int nchunks;
  const float* bias;
  float* Y; int ldy; int ygroup, ygstride;
  const float* Res; int ldr;
  int act;
  int M, N, K;
  uint32_t tmem_cols;
  const float *ln_w, *ln_b; float* Y2; int ldy2;      // optional fused LayerNorm(32) of the output rows -> Y2 (N == 32 only)
  // backward use (dX = dY . W): logical column k of A lives at (k / agroup) * agstride + k % agroup (agroup = 0: identity; agroup % 4 == 0);
  // the result is forced to 0 where Mask <= 0 (the ReLU of the layer whose input gradient this is), applied after the residual add
  int agroup, agstride;
  const float* Mask; int ldm;
  // sparse-convolution use (sparse_encoder.cu): the layer is Y = A_virtual . W_flat^T with K = 27 * kin and A_virtual[r][o * kin + k] =
  // A[rowtab[r * 27 + o]][k] (zero where the table holds -1): the A rows of chunk c are gathered through the neighbour table.  Mdev: the row
  // count lives on the device (the grid is sized by its upper bound M).  csplit > 0: split-K -- blockIdx.y handles the chunks
  // [blockIdx.y * csplit, ...) and writes its partial tile to Y + blockIdx.y * ysplit (summed in split order by the caller: deterministic)
  const int* rowtab; int kin;
  const int* Mdev;
  int csplit; size_t ysplit;
};

// Padded K-direction stride of the A operand: 2048 B of data + 16 B so that the 8 lanes of a quarter-warp that write 8
// different core-matrix columns of one row land in 8 different 16-byte bank groups (conflict-free STS.128).
constexpr uint32_t kALbo = 2064;

template <int PREC>
__global__ void __launch_bounds__(256) k_umma_linear(const UmmaArgs g) {
  constexpr int KC = (PREC >= 3) ? 32 : 64;          // k per shared-memory chunk
  constexpr int NKG = KC / 4;                        // core-matrix columns per chunk
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128;
  const int Meff = g.Mdev ? min(g.M, *g.Mdev) : g.M;
  if (m0 >= Meff) return;                            // (whole CTA, before any barrier or TMEM allocation)
  const uint32_t a_bytes = NKG * kALbo, w_bytes = (uint32_t)NKG * g.Np * 16u;
  float* A_hi = reinterpret_cast<float*>(smem);
  float* W_hi = reinterpret_cast<float*>(smem + a_bytes);
  float* A_lo = reinterpret_cast<float*>(smem + a_bytes + w_bytes);
  float* W_lo = reinterpret_cast<float*>(smem + 2 * a_bytes + w_bytes);
  float* stage = reinterpret_cast<float*>(smem);     // output tile, aliases the operands once all MMAs have completed
  const int sstride = g.Np + 4;

  if (tid == 0) { umma::mbar_init(&mma_bar, 1); umma::fence_mbar_init(); }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, g.tmem_cols);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  const uint32_t idesc = umma::make_idesc_tf32(128, g.Np);
  const int nchunks_all = (g.K + KC - 1) / KC;
  const int c_begin = g.csplit > 0 ? (int)blockIdx.y * g.csplit : 0;
  const int nchunks = g.csplit > 0 ? min(nchunks_all, c_begin + g.csplit) : nchunks_all;      // this CTA's chunks: [c_begin, nchunks)
  float* const Yout = g.Y + (size_t)blockIdx.y * g.ysplit;
  const int kgl = tid & 7, rsub = tid >> 3;          // loader mapping: 8 lanes = 8 consecutive float4 of one row (128 B)
  uint32_t parity = 0;
  // A chunk c: coalesced global reads -> registers (issued while chunk c - 1's MMAs run) -> tf32 hi (/lo) -> canonical smem
  constexpr int NPASS = NKG / 8;
  float4 pre[NPASS][4];
  auto load_a = [&](int c) {
    const int k0 = c * KC;
#pragma unroll
    for (int ps = 0; ps < NPASS; ++ps) {
      const int k = k0 + (ps * 8 + kgl) * 4;
#pragma unroll
      for (int rq = 0; rq < 4; ++rq) {
        const int row = rq * 32 + rsub;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.rowtab) {
          // gathered row: chunk -> kernel offset o and channel offset inside the neighbour's feature row (kin % KC == 0)
          const int o = k0 / g.kin;
          const int j = (m0 + row < Meff) ? __ldg(g.rowtab + (size_t)(m0 + row) * 27 + o) : -1;
          if (j >= 0) v = __ldg(reinterpret_cast<const float4*>(g.A + (size_t)j * g.lda + (k - o * g.kin)));
        } else if (m0 + row < Meff && k < g.K) {
          const float* ap = g.A + (size_t)(m0 + row) * g.lda + (g.agroup ? (k / g.agroup) * g.agstride + (k % g.agroup) : k);
          if (k + 3 < g.K) v = *reinterpret_cast<const float4*>(ap);
          else { v.x = ap[0]; if (k + 1 < g.K) v.y = ap[1]; if (k + 2 < g.K) v.z = ap[2]; }
        }
        pre[ps][rq] = v;
      }
    }
  };
  load_a(c_begin);
  for (int c = c_begin; c < nchunks; ++c) {
    const int k0 = c * KC;
    const int used_kg = min(NKG, (g.K - k0 + 3) / 4);
    const int mma_steps = (used_kg + 1) / 2;         // MMA K = 8 = two core-matrix columns
    if (c > c_begin) { umma::mbar_wait(&mma_bar, parity); parity ^= 1; }
    {
#pragma unroll
      for (int ps = 0; ps < NPASS; ++ps) {
        const int kg = ps * 8 + kgl;
        if (kg < 2 * mma_steps) {
#pragma unroll
          for (int rq = 0; rq < 4; ++rq) {
            const int row = rq * 32 + rsub;
            const float4 v = pre[ps][rq];
            const float4 h = make_float4(umma::to_tf32(v.x), umma::to_tf32(v.y), umma::to_tf32(v.z), umma::to_tf32(v.w));
            *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(A_hi) + kg * kALbo + row * 16) = h;
            if (PREC == 3) {
              const float4 l = make_float4(umma::to_tf32(v.x - h.x), umma::to_tf32(v.y - h.y), umma::to_tf32(v.z - h.z),
                                           umma::to_tf32(v.w - h.w));
              *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(A_lo) + kg * kALbo + row * 16) = l;
            }
          }
        }
      }
    }
    if (PREC == 4) {
      // experiment: A_lo goes to TENSOR MEMORY (lane = row, column = k), written row-per-thread with tcgen05.st
      const int q = warp & 3, hsel = warp >> 2;
      const int row = 32 * q + lane;
      for (int kg2 = hsel * (NKG / 2); kg2 < (hsel + 1) * (NKG / 2); kg2 += 2) {
        uint32_t lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = k0 + kg2 * 4 + e;
          const float x = (m0 + row < Meff && k < g.K) ? g.A[(size_t)(m0 + row) * g.lda + k] : 0.f;
          lo[e] = __float_as_uint(umma::to_tf32(x - umma::to_tf32(x)));
        }
        umma::tmem_st8(tmem_base + ((uint32_t)(32 * q) << 16) + 256u + (uint32_t)(kg2 * 4), lo);
      }
      umma::tmem_st_wait();
      umma::tc_fence_before_sync();
    }
    // ---- W chunk: contiguous pre-packed canonical block (only the core-matrix columns the MMAs will read) ----
    {
      const int n4 = 2 * mma_steps * g.Np;
      const float4* src_hi = reinterpret_cast<const float4*>(g.Whi) + (size_t)c * NKG * g.Np;
      float4* dst_hi = reinterpret_cast<float4*>(W_hi);
      for (int i = tid; i < n4; i += 256) dst_hi[i] = __ldg(src_hi + i);
      if (PREC >= 3) {
        const float4* src_lo = reinterpret_cast<const float4*>(g.Wlo) + (size_t)c * NKG * g.Np;
        float4* dst_lo = reinterpret_cast<float4*>(W_lo);
        for (int i = tid; i < n4; i += 256) dst_lo[i] = __ldg(src_lo + i);
      }
    }
    umma::fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {                                   // converged issuer warp, one elected lane issues
      umma::tc_fence_after_sync();
      const uint32_t a_hi_s = umma::smem_u32(A_hi), w_hi_s = umma::smem_u32(W_hi);
      const uint32_t a_lo_s = umma::smem_u32(A_lo), w_lo_s = umma::smem_u32(W_lo);
      const uint32_t w_lbo = (uint32_t)g.Np * 16u;
      for (int s = 0; s < mma_steps; ++s) {
        const uint32_t a_off = (uint32_t)s * 2u * kALbo, w_off = (uint32_t)s * 2u * w_lbo;
        const uint64_t ah = umma::make_smem_desc(a_hi_s + a_off, kALbo, 128u);
        const uint64_t wh = umma::make_smem_desc(w_hi_s + w_off, w_lbo, 128u);
        const uint32_t first = (c == c_begin && s == 0) ? 0u : 1u;
        if (PREC == 3) {
          const uint64_t al = umma::make_smem_desc(a_lo_s + a_off, kALbo, 128u);
          const uint64_t wl = umma::make_smem_desc(w_lo_s + w_off, w_lbo, 128u);
          umma::mma_tf32_ss_w(tmem_base, al, wh, idesc, first);          // small terms first
          umma::mma_tf32_ss_w(tmem_base, ah, wl, idesc, 1u);
          umma::mma_tf32_ss_w(tmem_base, ah, wh, idesc, 1u);
        } else if (PREC == 4) {
          const uint64_t wl = umma::make_smem_desc(w_lo_s + w_off, w_lbo, 128u);
          umma::mma_tf32_ts_w(tmem_base, tmem_base + 256u + (uint32_t)(s * 8), wh, idesc, first);
          umma::mma_tf32_ss_w(tmem_base, ah, wl, idesc, 1u);
          umma::mma_tf32_ss_w(tmem_base, ah, wh, idesc, 1u);
        } else {
          umma::mma_tf32_ss_w(tmem_base, ah, wh, idesc, first);
        }
      }
      umma::mma_commit_w(&mma_bar);
      __syncwarp();
    }
    if (c + 1 < nchunks) load_a(c + 1);
  }
  umma::mbar_wait(&mma_bar, parity);
  umma::tc_fence_after_sync();

  // ---- epilogue 1: TMEM -> registers -> bias / activation -> staging tile in smem.  Warp w owns lane quarter (w & 3). ----
  {
    const int q = warp & 3, hsel = warp >> 2;
    const int row = 32 * q + lane;
    const int ncol_half = g.Np / 2;
    for (int j = 0; j < ncol_half / 8; ++j) {
      const int c0 = hsel * ncol_half + 8 * j;
      uint32_t v[8];
      umma::tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
      umma::tmem_ld_wait();
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = c0 + i;
        x[i] = __uint_as_float(v[i]) + ((g.bias && n < g.N) ? __ldg(g.bias + n) : 0.f);
        if (g.act == 1) x[i] = fmaxf(x[i], 0.f);
        else if (g.act == 2) x[i] = 0.5f * x[i] * (1.f + erff(x[i] * 0.70710678118654752440f));
      }
      float4* dst = reinterpret_cast<float4*>(stage + (size_t)row * sstride + c0);
      dst[0] = make_float4(x[0], x[1], x[2], x[3]);
      dst[1] = make_float4(x[4], x[5], x[6], x[7]);
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, g.tmem_cols);
  // ---- epilogue 2: coalesced copy-out (+ residual, ReLU mask), consecutive lanes -> consecutive columns of a row.  Four tiles of 256
  //      float4 per pass: the residual / mask loads of a pass are all issued before anything consumes them (the first version read its
  //      mask with four dependent scalar loads per element group: 40 % of the kernel's stall samples in ncu) ----
  {
    const int n4 = (g.N + 3) / 4;                    // forward layers: N % 16 == 0; backward dX layers: N = 71 / 187 write one zero pad column
    const int total = 128 * n4;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(Yout) & 15) == 0) && (g.ldy % 4 == 0) && (g.ygroup % 4 == 0) && (g.ygstride % 4 == 0);
    const bool mvec_ok = g.Mask && ((reinterpret_cast<uintptr_t>(g.Mask) & 15) == 0) && (g.ldm % 4 == 0);
    for (int idx0 = tid; idx0 < total; idx0 += 1024) {
      float4 v[4], rr[4], mk[4];
      int mrow[4], ncol[4];
      bool valid[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = idx0 + 256 * u;
        const int row = idx / n4, c4 = idx - row * n4;
        mrow[u] = m0 + row; ncol[u] = c4 * 4;
        valid[u] = idx < total && mrow[u] < Meff;
        rr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        mk[u] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (idx < total) v[u] = *reinterpret_cast<const float4*>(stage + (size_t)row * sstride + ncol[u]);
        else v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.Res && valid[u]) rr[u] = *reinterpret_cast<const float4*>(g.Res + (size_t)mrow[u] * g.ldr + ncol[u]);
        if (g.Mask && valid[u]) {
          const float* mp = g.Mask + (size_t)mrow[u] * g.ldm + ncol[u];
          if (mvec_ok && ncol[u] + 3 < g.N) mk[u] = __ldg(reinterpret_cast<const float4*>(mp));
          else {
            mk[u].x = __ldg(mp);
            if (ncol[u] + 1 < g.N) mk[u].y = __ldg(mp + 1);
            if (ncol[u] + 2 < g.N) mk[u].z = __ldg(mp + 2);
            if (ncol[u] + 3 < g.N) mk[u].w = __ldg(mp + 3);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = ncol[u], m = mrow[u];
        float4 x = v[u];
        x.x += rr[u].x; x.y += rr[u].y; x.z += rr[u].z; x.w += rr[u].w;
        if (g.ln_w) {
          // LayerNorm over the 32 outputs of this row: the row lives in 8 consecutive lanes (n4 == 8, one full pass)    renderer.py:931
          float sum = x.x + x.y + x.z + x.w;
          sum += __shfl_xor_sync(0xffffffffu, sum, 1); sum += __shfl_xor_sync(0xffffffffu, sum, 2); sum += __shfl_xor_sync(0xffffffffu, sum, 4);
          const float mean = sum * (1.f / 32.f);
          const float d0 = x.x - mean, d1 = x.y - mean, d2 = x.z - mean, d3 = x.w - mean;
          float sq = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          sq += __shfl_xor_sync(0xffffffffu, sq, 1); sq += __shfl_xor_sync(0xffffffffu, sq, 2); sq += __shfl_xor_sync(0xffffffffu, sq, 4);
          const float rstd = rsqrtf(sq * (1.f / 32.f) + 1e-5f);
          if (valid[u]) {
            const float4 o = make_float4(d0 * rstd * g.ln_w[n] + g.ln_b[n], d1 * rstd * g.ln_w[n + 1] + g.ln_b[n + 1],
                                         d2 * rstd * g.ln_w[n + 2] + g.ln_b[n + 2], d3 * rstd * g.ln_w[n + 3] + g.ln_b[n + 3]);
            *reinterpret_cast<float4*>(g.Y2 + (size_t)m * g.ldy2 + n) = o;
          }
        }
        if (!valid[u]) continue;
        if (!(mk[u].x > 0.f)) x.x = 0.f;
        if (!(mk[u].y > 0.f)) x.y = 0.f;
        if (!(mk[u].z > 0.f)) x.z = 0.f;
        if (!(mk[u].w > 0.f)) x.w = 0.f;
        const int col = g.ygroup ? (n / g.ygroup) * g.ygstride + (n % g.ygroup) : n;
        float* yp = Yout + (size_t)m * g.ldy + col;
        if (vec_ok) *reinterpret_cast<float4*>(yp) = x;
        else { yp[0] = x.x; yp[1] = x.y; yp[2] = x.z; yp[3] = x.w; }
      }
    }
  }
}

static inline int round_up_i(int a, int b) { return (a + b - 1) / b * b; }

size_t canonical_weight_floats() {
  const int dims[16][2] = {{96, 192}, {32, 96}, {144, 32}, {32, 48}, {32, 32}, {32, 32}, {128, 71}, {128, 128}, {128, 128},
                           {128, 128}, {128, 128}, {128, 199}, {128, 128}, {128, 128}, {128, 128}, {64, 187}};
  size_t t = 0;
  for (int i = 0; i < 16; ++i) t += (size_t)round_up_i(dims[i][1], kKC) * round_up_i(dims[i][0], 16);
  return 2 * t;      // hi + lo
}

int run_pack_canonical(const SherfWeights& w, float* base, CanonWeights& cw, cudaStream_t st) {
  CanonJobs jobs;
  jobs.n = 0;
  float* cur = base;
  auto add = [&](CanonLayer& L, const float* W, const float* b, int N, int K) {
    L.N = N; L.K = K; L.Np = round_up_i(N, 16); L.nchunks = round_up_i(K, kKC) / kKC; L.bias = b;
    const size_t sz = (size_t)L.nchunks * kKC * L.Np;
    L.hi = cur; L.lo = cur + sz;
    CanonJob& j = jobs.j[jobs.n++];
    j.w = W; j.hi = cur; j.lo = cur + sz; j.N = N; j.K = K; j.Np = L.Np; j.nchunks = L.nchunks; j.ldw = 0; j.trans = 0;
    cur += 2 * sz;
  };
  add(cw.proj, w.proj_w, w.proj_b, 96, 192);
  add(cw.reproj, w.reproj_w, w.reproj_b, 32, 96);
  add(cw.qkv, w.qkv_w, nullptr, 144, 32);
  add(cw.attn_out, w.attn_out_w, w.attn_out_b, 32, 48);
  add(cw.ff1, w.ff1_w, w.ff1_b, 32, 32);
  add(cw.ff2, w.ff2_w, w.ff2_b, 32, 32);
  const int ptsK[8] = {71, 128, 128, 128, 128, 199, 128, 128};
  for (int i = 0; i < 8; ++i) add(cw.pts[i], w.pts_w[i], w.pts_b[i], 128, ptsK[i]);
  add(cw.feature, w.feature_w, w.feature_b, 128, 128);
  add(cw.views, w.views_w, w.views_b, 64, 187);
  if (!g_pack_plan_only) {
    k_pack_canonical<<<dim3(16, jobs.n), 256, 0, st>>>(jobs);
    SHERF_LAUNCH_CHECK();
  }
  return SHERF_OK;
}

// extra operands of the backward's dX launches (set around one launch_umma_linear call by launch_umma_dx; zero otherwise)
struct UmmaEx { int agroup, agstride; const float* Mask; int ldm; const int* rowtab; int kin; const int* Mdev; int csplit; size_t ysplit; int nsplit; };
static thread_local UmmaEx g_umma_ex = {0, 0, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0};

int launch_umma_linear(int prec, const CanonLayer& L, const float* A, int lda, float* Y, int ldy, int M, int act, cudaStream_t st,
                       const float* Res, int ldr, int ygroup, int ygstride, const float* ln_w, const float* ln_b, float* Y2, int ldy2) {
  UmmaArgs g;
  g.A = A; g.lda = lda; g.Whi = L.hi; g.Wlo = L.lo; g.Np = L.Np; g.nchunks = L.nchunks; g.bias = L.bias;
  g.Y = Y; g.ldy = ldy; g.ygroup = ygroup; g.ygstride = ygstride; g.Res = Res; g.ldr = ldr; g.act = act;
  g.M = M; g.N = L.N; g.K = L.K;
  g.ln_w = (L.N == 32) ? ln_w : nullptr; g.ln_b = ln_b; g.Y2 = Y2; g.ldy2 = ldy2;
  g.agroup = g_umma_ex.agroup; g.agstride = g_umma_ex.agstride; g.Mask = g_umma_ex.Mask; g.ldm = g_umma_ex.ldm;
  g.rowtab = g_umma_ex.rowtab; g.kin = g_umma_ex.kin; g.Mdev = g_umma_ex.Mdev; g.csplit = g_umma_ex.csplit; g.ysplit = g_umma_ex.ysplit;
  const int gy = g_umma_ex.nsplit > 0 ? g_umma_ex.nsplit : 1;
  uint32_t cols = 32;
  while ((int)cols < L.Np) cols <<= 1;
  g.tmem_cols = cols;
  const size_t nkg = prec >= 3 ? 8 : 16;
  const size_t operands = (size_t)(prec >= 3 ? 2 : 1) * (nkg * kALbo + nkg * (size_t)L.Np * 16);
  const size_t staging = (size_t)128 * (L.Np + 4) * sizeof(float);
  const size_t smem = operands > staging ? operands : staging;
  static bool attr_done = false;
  if (!attr_done) {
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_umma_linear<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_umma_linear<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  if (prec == 4) {
    static bool a4 = false;
    if (!a4) { SHERF_CUDA_OK(cudaFuncSetAttribute(k_umma_linear<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); a4 = true; }
    g.tmem_cols = 512;
    k_umma_linear<4><<<dim3(ceil_div(M, 128), gy), 256, smem, st>>>(g);
  } else if (prec == 3) k_umma_linear<3><<<dim3(ceil_div(M, 128), gy), 256, smem, st>>>(g);
  else k_umma_linear<1><<<dim3(ceil_div(M, 128), gy), 256, smem, st>>>(g);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}


// Single linear layer with freshly packed weights (tests / diagnostics).  wscratch: >= 8*272*272 floats.
__global__ void k_pack_plain(const float* __restrict__ w, float* __restrict__ wt, int N, int K, int kp, int np) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kp * np; i += gridDim.x * blockDim.x) {
    const int k = i / np, n = i - k * np;
    wt[i] = (k < K && n < N) ? w[(size_t)n * K + k] : 0.f;
  }
}

int run_debug_linear(int prec, const float* A, int lda, const float* W, const float* bias, float* Y, int ldy, int M, int N, int K,
                     int act, float* wscratch, cudaStream_t st) {
  if (prec == SHERF_MLP_FP32) {
    PackedLayer L;
    L.K = K; L.N = N; L.kp = round_up_i(K, 16); L.np = round_up_i(N, 32); L.wt = wscratch; L.bias = bias;
    k_pack_plain<<<64, 256, 0, st>>>(W, wscratch, N, K, L.kp, L.np);
    SHERF_LAUNCH_CHECK();
    return launch_simt_linear(L, A, lda, Y, ldy, M, act, st, nullptr, 0, 0, 0);
  }
  if (N % 16 != 0) { set_error("tensor-core linear needs N %% 16 == 0 (got %d)", N); return SHERF_E_INVALID; }
  CanonJobs jobs;
  CanonLayer L;
  L.N = N; L.K = K; L.Np = round_up_i(N, 16); L.nchunks = round_up_i(K, kKC) / kKC; L.bias = bias;
  const size_t sz = (size_t)L.nchunks * kKC * L.Np;
  L.hi = wscratch; L.lo = wscratch + sz;
  jobs.n = 1;
  jobs.j[0].w = W; jobs.j[0].hi = wscratch; jobs.j[0].lo = wscratch + sz; jobs.j[0].N = N; jobs.j[0].K = K; jobs.j[0].Np = L.Np;
  jobs.j[0].nchunks = L.nchunks; jobs.j[0].ldw = 0; jobs.j[0].trans = 0;
  k_pack_canonical<<<dim3(16, 1), 256, 0, st>>>(jobs);
  SHERF_LAUNCH_CHECK();
  return launch_umma_linear(prec == 99 ? 4 : (prec == SHERF_MLP_TF32X3 ? 3 : 1), L, A, lda, Y, ldy, M, act, st, nullptr, 0, 0, 0);
}

}  // namespace sherf

// ------------------------------------------------------------------------------------------------- backward: dX = dY . W on the tensor cores
// The same kernel with the weight matrix packed transposed: "output feature" n = the layer's INPUT feature, reduction k = its OUTPUT feature.
namespace sherf {

size_t canonical_bwd_weight_floats() {
  // (in, out) of the 17 dX products of backward.cu
  const int dims[17][2] = {{187, 64}, {128, 128}, {71, 128}, {128, 128}, {128, 128}, {128, 128}, {128, 128}, {128, 128}, {128, 128}, {128, 128},
                           {71, 128}, {32, 32}, {32, 32}, {48, 32}, {32, 144}, {96, 32}, {192, 96}};
  size_t t = 0;
  for (int i = 0; i < 17; ++i) t += (size_t)round_up_i(dims[i][1], kKC) * round_up_i(dims[i][0], 16);
  return 2 * t;
}

int run_pack_canonical_bwd(const SherfWeights& w, float* base, CanonBwdWeights& cb, cudaStream_t st) {
  CanonJobs jobs;
  jobs.n = 0;
  float* cur = base;
  // W: [out][ldw] PyTorch layout, columns col0 .. col0 + in - 1 of it
  auto add = [&](CanonLayer& L, const float* W, int out, int in, int ldw) {
    L.N = in; L.K = out; L.Np = round_up_i(in, 16); L.nchunks = round_up_i(out, kKC) / kKC; L.bias = nullptr;
    const size_t sz = (size_t)L.nchunks * kKC * L.Np;
    L.hi = cur; L.lo = cur + sz;
    CanonJob& j = jobs.j[jobs.n++];
    j.w = W; j.hi = cur; j.lo = cur + sz; j.N = in; j.K = out; j.Np = L.Np; j.nchunks = L.nchunks; j.ldw = ldw; j.trans = 1;
    cur += 2 * sz;
  };
  add(cb.views, w.views_w, 64, 187, 187);
  add(cb.feature, w.feature_w, 128, 128, 128);
  add(cb.pts[0], w.pts_w[0], 128, 71, 71);
  for (int i = 1; i < 8; ++i) {
    if (i == 5) add(cb.pts[5], w.pts_w[5] + 71, 128, 128, 199);      // the h4 half of cat([x, h4]) (triplane.py:299-300)
    else add(cb.pts[i], w.pts_w[i], 128, 128, 128);
  }
  add(cb.pts5x, w.pts_w[5], 128, 71, 199);                           // the x half
  add(cb.ff2, w.ff2_w, 32, 32, 32);
  add(cb.ff1, w.ff1_w, 32, 32, 32);
  add(cb.attn_out, w.attn_out_w, 32, 48, 48);
  add(cb.qkv, w.qkv_w, 144, 32, 32);
  add(cb.reproj, w.reproj_w, 32, 96, 96);
  add(cb.proj, w.proj_w, 96, 192, 192);
  k_pack_canonical<<<dim3(16, jobs.n), 256, 0, st>>>(jobs);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

int launch_umma_dx(const CanonLayer& L, const float* dY, int lda, float* dX, int ldx, int M, cudaStream_t st, const float* Mask, int ldm,
                   int accum, int agroup, int agstride) {
  g_umma_ex = UmmaEx{agroup, agstride, Mask, ldm, nullptr, 0, nullptr, 0, 0, 0};
  const int rc = launch_umma_linear(3, L, dY, lda, dX, ldx, M, 0, st, accum ? dX : nullptr, ldx, 0, 0);
  g_umma_ex = UmmaEx{0, 0, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0};
  return rc;
}

// canonical pack of one sparse-convolution weight [c_out][27][c_in] (KRSC).  mode 0: forward, N = c_out, K = 27 c_in (the flattened weight);
// mode 2 / 3: input gradient (N = c_in, K = 27 c_out), mirrored offsets (SubMConv3d) / plain offsets (strided convolution)
size_t spconv_canon_floats() { return (size_t)2 * round_up_i(27 * 96, kKC) * 96; }
int run_pack_spconv(const float* W, int cout, int cin, int mode, float* buf, CanonLayer& L, cudaStream_t st) {
  CanonJobs jobs;
  jobs.n = 1;
  const int N = mode == 0 ? cout : cin, K = 27 * (mode == 0 ? cin : cout);
  L.N = N; L.K = K; L.Np = round_up_i(N, 16); L.nchunks = round_up_i(K, kKC) / kKC; L.bias = nullptr;
  const size_t sz = (size_t)L.nchunks * kKC * L.Np;
  L.hi = buf; L.lo = buf + sz;
  CanonJob& j = jobs.j[0];
  j.w = W; j.hi = buf; j.lo = buf + sz; j.N = N; j.K = K; j.Np = L.Np; j.nchunks = L.nchunks; j.ldw = 0; j.trans = mode;
  k_pack_canonical<<<dim3(64, 1), 256, 0, st>>>(jobs);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

// Sparse convolution as a gathered linear layer (sparse_encoder.cu): Y_part[s][r][n] = sum over the chunks of split s of
// X[rowtab[r][o]][k] W[n][o][k].  L: canonical pack of W viewed as [N][27 * kin]; Mcap sizes the grid, *Mdev is the row count.
int launch_umma_spconv(const CanonLayer& L, const float* X, int kin, const int* rowtab, const int* Mdev, int Mcap, float* Ypart, int nsplit,
                       cudaStream_t st) {
  const int nchunks = L.K / 32;                      // 27 * kin / 32 chunks of the 3xTF32 kernel
  const int csplit = ceil_div(nchunks, nsplit);
  g_umma_ex = UmmaEx{0, 0, nullptr, 0, rowtab, kin, Mdev, csplit, (size_t)Mcap * L.N, nsplit};
  const int rc = launch_umma_linear(3, L, X, kin, Ypart, L.N, Mcap, 0, st, nullptr, 0, 0, 0);
  g_umma_ex = UmmaEx{0, 0, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0};
  return rc;
}

}  // namespace sherf
