// Fused 3-token transformer layer (renderer.py:920-993) + decoder-input assembly (renderer.py:432, triplane.py:293,308) on the
// tensor cores with bf16 SPLIT PRODUCTS, TWO co-resident CTAs per SM.
//
// Same algorithm and phase structure as xformer_fused.cu (rows of every MMA = 128 points, the three tokens of a point live in
// column blocks of one TMEM lane, so the 3x3 attention / LayerNorm / residual / GELU epilogues are thread-local):
//   phase            MMA (M=128 points)                         epilogue (thread = (point, query token t in {0,1}))
//   (load)                                                     LayerNorm-1 of the point's three tokens -> LN1 operands
//   qkv, head h=0..2 D3[:, j*48:+48] = LN1_j * Wqkv_h^T  j=0..2   softmax(q_t k_j^T / 4) v_j -> ATT_t[:, h*16:+16]
//   to_out           D4[:, t*32:+32] = ATT_t * Wo^T               + bias + tok_t -> tok2_t ; LayerNorm -> LN2_t
//   ff1              D5[:, t*32:+32] = LN2_t * W1^T               GELU(. + b1) -> G_t
//   ff2              D4[:, t*32:+32] = G_t * W2^T                 + b2 + tok2_t -> tok3_t -> packed decoder inputs
// What changed against the 3xTF32 kernel, and why (profiles/r1_ab: it was a 7-hand-off latency chain with ONE tile in flight
// per SM, tensor pipe 15 %):
//   * arithmetic: a = a_hi + a_lo, w = w_hi + w_lo in bf16; a_hi*w_hi + a_lo*w_hi + a_hi*w_lo on tcgen05 kind::f16 with fp32
//     accumulation (the decoder's scheme, 16 significand bits per operand).  Operand bytes halve: hi parts in shared memory
//     (49.5 KB), lo parts in tensor memory, resident weights 32 KB -> 82 KB of shared memory and 240 TMEM columns per CTA
//     (D4 / D5 alias the qkv accumulator), so TWO CTAs fit an SM and two tiles are in flight without an in-kernel ping-pong;
//   * the first LayerNorm is computed here from the tokens (the fusion kernel writes only the tokens);
//   * attention epilogue restructured to hold q, one k_j / v_j at a time (<= 112 registers per thread for 2 CTAs x 288 threads).
// Warp roles per CTA: warps 0-7 = tile loader + epilogues, warp 8 = converged MMA issuer.
#include "common.cuh"
#include "stages.cuh"
#include "umma.cuh"
#include <cuda_bf16.h>

namespace sherf {

namespace xb {
constexpr uint32_t kLbo = 2064;                     // bytes between core-matrix columns (8 k-elements = 16 B per row) of an A operand
// shared memory map (bytes)
constexpr uint32_t kBuf1 = 0;                       // LN1 (3 tokens x 4 kg) -> LN2 (2 x 4 kg at token*4)
constexpr uint32_t kBuf2 = 12 * kLbo;               // ATT (2 tokens x 6 kg) -> G (2 x 4 kg at token*6)
constexpr uint32_t kW = 24 * kLbo;                  // resident weights: hi block then lo block
// weight blocks (bf16 elements within the hi part; the lo part follows at +kWElems): element (kg, n, e) = W_block[n][kg*8 + e]
constexpr int kWqkv = 0;                            // 3 heads x [4 kg][48 rows][8]
constexpr int kWo = 3 * 4 * 48 * 8;                 // [6 kg][32][8]
constexpr int kW1 = kWo + 6 * 32 * 8;               // [4 kg][32][8]
constexpr int kW2 = kW1 + 4 * 32 * 8;
constexpr int kWElems = kW2 + 4 * 32 * 8;           // 8192
constexpr uint32_t kSmemBytes = kW + 2 * kWElems * 2;
// tensor memory map (columns, 256 allocated)
constexpr uint32_t kD3 = 0, kD4 = 0, kD5 = 64, kLo1 = 144, kLo2 = 192;     // LN1_lo / LN2_lo at kLo1 (16 columns per token), ATT_lo / G_lo at kLo2 (24 per token)
}  // namespace xb

struct XbArgs {
  const float *tok, *geo;              // [3np][32] tokens (conv1d_reprojection output), [np][8] can / cdir
  const unsigned char* wblob;          // canonical bf16 hi | lo
  const float *ln1_w, *ln1_b, *bo, *ln_w, *ln_b, *b1, *b2;
  unsigned char *xp, *vp;              // packed bf16 hi/lo decoder-input tiles (decoder_pp.cu)
  float* dbg_tok; int64_t p0, dbg_max;
  int np;
  DevCount dc;
};

// erf for the GELU (renderer.py:936-947, nn.GELU() = 0.5 x (1 + erf(x / sqrt 2))): branch-free erf(|x|) = 1 - exp(-|x| q(|x|)), q a degree-7
// polynomial fitted on [0, 4] (erf rounds to 1 in fp32 beyond 3.92), sign restored.  Maximum absolute error 1.6e-7 (fp32 evaluation, checked
// against scipy over [0, 6] in steps of 2e-6: tests/test_oracle.py) -- two ulps of 1.0, the size of the rounding of `1 + erf` itself -- at
// 22 instructions; CUDA's two-branch erff cost 38 per call with both branches executed by every warp, a quarter of this kernel's instructions.
__device__ __forceinline__ float xb_erf(float x) {
  const float t = fminf(fabsf(x), 4.0f);
  float q = 2.5281295165768825e-05f;
  q = fmaf(q, t, -0.00025927156093530357f);
  q = fmaf(q, t, 0.0008752066642045975f);
  q = fmaf(q, t, 0.0007889552507549524f);
  q = fmaf(q, t, -0.01980074681341648f);
  q = fmaf(q, t, 0.10301736742258072f);
  q = fmaf(q, t, 0.6365770697593689f);
  q = fmaf(q, t, 1.1283817291259766f);
  return copysignf(1.0f - expf(-t * q), x);
}

__device__ __forceinline__ void xb_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(288, 2) k_xformer_bf16(const XbArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* buf1 = smem + xb::kBuf1;
  unsigned char* buf2 = smem + xb::kBuf2;
  __shared__ __align__(8) uint64_t acc_bar, a_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_ln1w[32], s_ln1b[32], s_bo[32], s_lnw[32], s_lnb[32], s_b1[32], s_b2[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) { umma::mbar_init(&acc_bar, 1); umma::mbar_init(&a_bar, 256); umma::fence_mbar_init(); }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  for (int i = tid; i < (int)(2 * xb::kWElems * 2 / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(smem + xb::kW)[i] = __ldg(reinterpret_cast<const uint4*>(a.wblob) + i);
  if (tid < 32) {
    s_ln1w[tid] = a.ln1_w[tid]; s_ln1b[tid] = a.ln1_b[tid]; s_bo[tid] = a.bo[tid]; s_lnw[tid] = a.ln_w[tid]; s_lnb[tid] = a.ln_b[tid];
    s_b1[tid] = a.b1[tid]; s_b2[tid] = a.b2[tid];
  }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int np = resolve_np(a.np, a.dc);
  const int ntiles = (np + 127) / 128;

  if (warp == 8) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    uint32_t par_a = 0;
    const uint32_t el = umma::elect_one();
    const uint32_t b1s = umma::smem_u32(buf1), b2s = umma::smem_u32(buf2);
    const uint32_t whs = umma::smem_u32(smem + xb::kW), wls = whs + xb::kWElems * 2;
    // one GEMM block: D[:, dcol:+N] = A[128 x 16*nks] * W[N x 16*nks]^T with split products
    auto gemm = [&](uint32_t a_hi_addr, uint32_t a_lo_col, int w_off_elems, int N, int nks, uint32_t dcol) {
      const uint32_t idesc = umma::make_idesc_bf16(128, N);
      const uint32_t w_lbo = (uint32_t)N * 16u;
      const uint64_t ah0 = umma::make_smem_desc(a_hi_addr, xb::kLbo, 128u);
      const uint64_t wh0 = umma::make_smem_desc(whs + (uint32_t)w_off_elems * 2u, w_lbo, 128u);
      const uint64_t wl0 = umma::make_smem_desc(wls + (uint32_t)w_off_elems * 2u, w_lbo, 128u);
      const uint64_t da = (uint64_t)((2u * xb::kLbo) >> 4), dw = (uint64_t)((2u * w_lbo) >> 4);
      for (int st = 0; st < nks; ++st) {
        umma::mma_bf16_ts_e(tmem_base + dcol, tmem_base + a_lo_col + (uint32_t)st * 8u, wh0 + (uint64_t)st * dw, idesc, st == 0 ? 0u : 1u, el);
        umma::mma_bf16_ss_e(tmem_base + dcol, ah0 + (uint64_t)st * da, wl0 + (uint64_t)st * dw, idesc, 1u, el);
        umma::mma_bf16_ss_e(tmem_base + dcol, ah0 + (uint64_t)st * da, wh0 + (uint64_t)st * dw, idesc, 1u, el);
      }
    };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int ph = 0; ph < 6; ++ph) {
        umma::mbar_wait(&a_bar, par_a);
        par_a ^= 1;
        umma::tc_fence_after_sync();
        if (ph < 3) {                       // qkv of head ph for the three tokens
          for (int j = 0; j < 3; ++j)
            gemm(b1s + (uint32_t)(j * 4) * xb::kLbo, xb::kLo1 + (uint32_t)(j * 16), xb::kWqkv + ph * 4 * 48 * 8, 48, 2, xb::kD3 + (uint32_t)(j * 48));
        } else if (ph == 3) {               // to_out for the two query tokens
          for (int t = 0; t < 2; ++t)
            gemm(b2s + (uint32_t)(t * 6) * xb::kLbo, xb::kLo2 + (uint32_t)(t * 24), xb::kWo, 32, 3, xb::kD4 + (uint32_t)(t * 32));
        } else if (ph == 4) {               // ff1
          for (int t = 0; t < 2; ++t)
            gemm(b1s + (uint32_t)(t * 4) * xb::kLbo, xb::kLo1 + (uint32_t)(t * 16), xb::kW1, 32, 2, xb::kD5 + (uint32_t)(t * 32));
        } else {                            // ff2
          for (int t = 0; t < 2; ++t)
            gemm(b2s + (uint32_t)(t * 6) * xb::kLbo, xb::kLo2 + (uint32_t)(t * 24), xb::kW2, 32, 2, xb::kD4 + (uint32_t)(t * 32));
        }
        umma::mma_commit_e(&acc_bar, el);
      }
    }
  } else {
    // ===================== tile loader + epilogues (warps 0-7): thread = (point row, query token t) =====================
    const int q = warp & 3, t = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t tb = tmem_base + ((uint32_t)(32 * q) << 16);
    uint32_t par_acc = 0;
    // 16 consecutive k-elements of this row: bf16 hi -> two 16-byte chunks (core-matrix columns kg0, kg0+1), bf16 lo -> 8 packed TMEM columns
    auto split_store = [&](unsigned char* buf, int kg0, uint32_t lo_col, const float (&v)[16]) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) umma::split_bf16x2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
      *reinterpret_cast<uint4*>(buf + (size_t)kg0 * xb::kLbo + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(buf + (size_t)(kg0 + 1) * xb::kLbo + row * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      umma::tmem_st8(tb + lo_col, lo);
    };
    auto layernorm32 = [&](const float (&x)[32], const float* __restrict__ gw, const float* __restrict__ gb, float (&y)[32]) {
      float mean = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) mean += x[i];
      mean *= (1.f / 32.f);
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) { const float d = x[i] - mean; var += d * d; }
      const float rstd = rsqrtf(var * (1.f / 32.f) + 1e-5f);
#pragma unroll
      for (int i = 0; i < 32; ++i) y[i] = (x[i] - mean) * rstd * gw[i] + gb[i];
    };
    auto load32 = [&](const float* src, bool ok, float (&x)[32]) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 f = ok ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        x[4 * i] = f.x; x[4 * i + 1] = f.y; x[4 * i + 2] = f.z; x[4 * i + 3] = f.w;
      }
    };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m = tile * 128 + row;
      const bool row_ok = m < np;
      // ---- LayerNorm-1 (renderer.py:931): token t fully by this thread, token 2's statistics by both threads of the row, each
      //      writing one half of its operand ----
      {
        float y[32], x1[32];
        load32(a.tok + (size_t)(m * 3 + t) * 32, row_ok, x1);
        layernorm32(x1, s_ln1w, s_ln1b, y);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = y[16 * half + i];
          split_store(buf1, t * 4 + half * 2, xb::kLo1 + (uint32_t)(t * 16 + half * 8), v);
        }
        float x2[32];
        load32(a.tok + (size_t)(m * 3 + 2) * 32, row_ok, x2);
        float v[16];
        {                                                      // statistics over the whole token, normalisation of this thread's half only
          float mean = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) mean += x2[i];
          mean *= (1.f / 32.f);
          float var = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) { const float d = x2[i] - mean; var += d * d; }
          const float rstd = rsqrtf(var * (1.f / 32.f) + 1e-5f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float xv = t == 0 ? x2[i] : x2[16 + i];
            v[i] = (xv - mean) * rstd * s_ln1w[16 * t + i] + s_ln1b[16 * t + i];
          }
        }
        split_store(buf1, 8 + t * 2, xb::kLo1 + (uint32_t)(32 + t * 8), v);
        umma::tmem_st_wait();
      }
      umma::fence_proxy_async_smem();
      umma::tc_fence_before_sync();
      xb_arrive(&a_bar);
      // pull the CTA's next tile (tokens, geometry, encodings) towards L2 while the tensor core works
      {
        auto pf = [](const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
        const int mn = m + (int)gridDim.x * 128;
        if (mn < np) {
          pf(a.tok + (size_t)(mn * 3 + t) * 32);
          if (t == 0) { pf(a.tok + (size_t)(mn * 3 + 2) * 32); pf(a.geo + (size_t)mn * 8); }
        }
      }
      // ---- three heads: attention of query token t over the three tokens ----
      for (int h = 0; h < 3; ++h) {
        umma::mbar_wait(&acc_bar, par_acc);
        par_acc ^= 1;
        umma::tc_fence_after_sync();
        uint32_t qv[16];
        umma::tmem_ld16(tb + xb::kD3 + (uint32_t)(t * 48), qv);
        float dots[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t kv[16];
          umma::tmem_ld16(tb + xb::kD3 + (uint32_t)(j * 48 + 16), kv);
          umma::tmem_ld_wait();
          float s = 0.f;
#pragma unroll
          for (int d = 0; d < 16; ++d) s += __uint_as_float(qv[d]) * __uint_as_float(kv[d]);
          dots[j] = s * 0.25f;                              // dim_head ** -0.5   (renderer.py:956,971)
        }
        const float mx = fmaxf(dots[0], fmaxf(dots[1], dots[2]));
        const float e0 = expf(dots[0] - mx), e1 = expf(dots[1] - mx), e2 = expf(dots[2] - mx);
        const float inv = 1.f / (e0 + e1 + e2);
        const float aw[3] = {e0 * inv, e1 * inv, e2 * inv};
        float att[16];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t vv[16];
          umma::tmem_ld16(tb + xb::kD3 + (uint32_t)(j * 48 + 32), vv);
          umma::tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < 16; ++d) att[d] = j == 0 ? aw[0] * __uint_as_float(vv[d]) : att[d] + aw[j] * __uint_as_float(vv[d]);
        }
        split_store(buf2, t * 6 + h * 2, xb::kLo2 + (uint32_t)(t * 24 + h * 8), att);
        umma::tmem_st_wait();
        umma::fence_proxy_async_smem();
        umma::tc_fence_before_sync();
        xb_arrive(&a_bar);
      }
      // ---- to_out + residual + LayerNorm-2 ----
      float tok2[32];
      {
        umma::mbar_wait(&acc_bar, par_acc);
        par_acc ^= 1;
        umma::tc_fence_after_sync();
        uint32_t o0[16], o1[16];
        umma::tmem_ld16(tb + xb::kD4 + (uint32_t)(t * 32), o0);
        umma::tmem_ld16(tb + xb::kD4 + (uint32_t)(t * 32 + 16), o1);
        load32(a.tok + (size_t)(m * 3 + t) * 32, row_ok, tok2);          // the residual: this thread's own token again (L1 / L2 hit)
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) { tok2[i] += __uint_as_float(o0[i]) + s_bo[i]; tok2[16 + i] += __uint_as_float(o1[i]) + s_bo[16 + i]; }
        float y[32];
        layernorm32(tok2, s_lnw, s_lnb, y);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = y[16 * half + i];
          split_store(buf1, t * 4 + half * 2, xb::kLo1 + (uint32_t)(t * 16 + half * 8), v);
        }
        umma::tmem_st_wait();
        umma::fence_proxy_async_smem();
        umma::tc_fence_before_sync();
        xb_arrive(&a_bar);
      }
      // ---- ff1 + GELU ----
      {
        umma::mbar_wait(&acc_bar, par_acc);
        par_acc ^= 1;
        umma::tc_fence_after_sync();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t o[16];
          umma::tmem_ld16(tb + xb::kD5 + (uint32_t)(t * 32 + 16 * half), o);
          umma::tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float x = __uint_as_float(o[i]) + s_b1[16 * half + i]; v[i] = 0.5f * x * (1.f + xb_erf(x * 0.70710678118654752440f)); }
          split_store(buf2, t * 6 + half * 2, xb::kLo2 + (uint32_t)(t * 24 + half * 8), v);
        }
        umma::tmem_st_wait();
        umma::fence_proxy_async_smem();
        umma::tc_fence_before_sync();
        xb_arrive(&a_bar);
      }
      // ---- ff2 + residual -> packed decoder inputs ----
      {
        const float g0 = row_ok ? a.geo[(size_t)m * 8 + 3 * t] : 0.f, g1 = row_ok ? a.geo[(size_t)m * 8 + 3 * t + 1] : 0.f,
                    g2 = row_ok ? a.geo[(size_t)m * 8 + 3 * t + 2] : 0.f;
        umma::mbar_wait(&acc_bar, par_acc);
        par_acc ^= 1;
        umma::tc_fence_after_sync();
        uint32_t o0[16], o1[16];
        umma::tmem_ld16(tb + xb::kD4 + (uint32_t)(t * 32), o0);
        umma::tmem_ld16(tb + xb::kD4 + (uint32_t)(t * 32 + 16), o1);
        umma::tmem_ld_wait();
        umma::tc_fence_before_sync();                      // the accumulator is drained: the next tile's MMAs may overwrite it
        float tok3[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) { tok3[i] = __uint_as_float(o0[i]) + s_b2[i] + tok2[i]; tok3[16 + i] = __uint_as_float(o1[i]) + s_b2[16 + i] + tok2[16 + i]; }
        // decoder inputs, one 8-column core-matrix row at a time (nothing but tok3 and the point's xyz stays live):
        //   t = 0: X = [can(3) | PE6(can)(36) | tok0(32)] = 71 -> 80 columns;  t = 1: V = [cdir(3) | PE4(cdir)(24) | tok1(32)] = 59 -> 64 columns
        // PositionalEncoding (renderer.py:900-916): pe[3*mm + c] = sin(phase(mm) + x_c * 2^(mm >> 1)), phase = 0 | pi/2, computed here
        // with torch.addcmul's separately rounded multiply and add (the round-1 k_point_pe pass and its 256 B per point are gone).
        auto col_value = [&](int col, int npe) -> float {              // `col` is a compile-time constant after unrolling
          if (col < 3) return col == 0 ? g0 : (col == 1 ? g1 : g2);
          if (col < 3 + npe) {
            const int o = col - 3, mm = o / 3, cc = o - 3 * mm;
            const float x = cc == 0 ? g0 : (cc == 1 ? g1 : g2);
            return sinf(__fadd_rn((mm & 1) ? kPi2 : 0.f, __fmul_rn(x, (float)(1 << (mm >> 1)))));
          }
          const int o = col - 3 - npe;
          return o < 32 ? tok3[o] : 0.f;
        };
        auto put8 = [&](unsigned char* tile_base, int kg, uint32_t lo_off, const float (&v8)[8]) {
          uint4 h, l;
          umma::split_bf16x2(v8[0], v8[1], h.x, l.x); umma::split_bf16x2(v8[2], v8[3], h.y, l.y);
          umma::split_bf16x2(v8[4], v8[5], h.z, l.z); umma::split_bf16x2(v8[6], v8[7], h.w, l.w);
          *reinterpret_cast<uint4*>(tile_base + (size_t)(kg * 128 + row) * 16) = h;
          *reinterpret_cast<uint4*>(tile_base + lo_off + (size_t)(kg * 128 + row) * 16) = l;
        };
        if (t == 0) {
          unsigned char* tb_ = a.xp + (size_t)tile * 40960;
#pragma unroll
          for (int kg = 0; kg < 10; ++kg) {
            float v8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v8[e] = row_ok ? col_value(8 * kg + e, 36) : 0.f;
            put8(tb_, kg, 20480u, v8);
          }
        } else {
          unsigned char* tb_ = a.vp + (size_t)tile * 32768;
#pragma unroll
          for (int kg = 0; kg < 8; ++kg) {
            float v8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v8[e] = row_ok ? col_value(8 * kg + e, 24) : 0.f;
            put8(tb_, kg, 16384u, v8);
          }
        }
        if (row_ok && a.dbg_tok && a.p0 + m < a.dbg_max) {
#pragma unroll
          for (int o = 0; o < 32; ++o) a.dbg_tok[(a.p0 + m) * 64 + t * 32 + o] = tok3[o];
        }
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------------------
// canonical bf16 weight blob: element (kg, n, e) of a block = W_block[n][kg*8 + e]; hi part [kWElems] then lo part [kWElems]
__global__ void k_pack_xformer_bf16(const float* __restrict__ wqkv, const float* __restrict__ wo, const float* __restrict__ w1,
                                    const float* __restrict__ w2, __nv_bfloat16* __restrict__ blob) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= xb::kWElems) return;
  float v;
  if (i < xb::kWo) {                                        // head h block: rows [q_h(16) | k_h(16) | v_h(16)] of to_qkv.weight [144][32]
    const int h = i / (4 * 48 * 8), r = i % (4 * 48 * 8);
    const int e = r & 7, n = (r >> 3) % 48, kg = (r >> 3) / 48;
    const int src_row = (n / 16) * 48 + h * 16 + (n % 16);  // q block at 0, k at 48, v at 96; head h at +16h   (renderer.py:968-969)
    v = wqkv[src_row * 32 + kg * 8 + e];
  } else if (i < xb::kW1) {
    const int r = i - xb::kWo;
    const int e = r & 7, n = (r >> 3) % 32, kg = (r >> 3) / 32;
    v = wo[n * 48 + kg * 8 + e];
  } else if (i < xb::kW2) {
    const int r = i - xb::kW1;
    const int e = r & 7, n = (r >> 3) % 32, kg = (r >> 3) / 32;
    v = w1[n * 32 + kg * 8 + e];
  } else {
    const int r = i - xb::kW2;
    const int e = r & 7, n = (r >> 3) % 32, kg = (r >> 3) / 32;
    v = w2[n * 32 + kg * 8 + e];
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  blob[i] = h;
  blob[xb::kWElems + i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

size_t xformer_bf16_blob_bytes() { return 2 * (size_t)xb::kWElems * 2; }

int run_pack_xformer_bf16(const SherfWeights& w, unsigned char* blob, cudaStream_t st) {
  k_pack_xformer_bf16<<<ceil_div(xb::kWElems, 256), 256, 0, st>>>(w.qkv_w, w.attn_out_w, w.ff1_w, w.ff2_w, reinterpret_cast<__nv_bfloat16*>(blob));
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

int run_xformer_bf16(const SherfWeights& w, const unsigned char* blob, const float* tok, const float* geo, int np, float* dbg_tok, int64_t p0,
                     int64_t dbg_max, cudaStream_t st, unsigned char* xp, unsigned char* vp, float* pe_buf, DevCount dc) {
  if (np <= 0) return SHERF_OK;
  (void)pe_buf;                                             // positional encodings are computed in the last epilogue
  XbArgs a;
  a.tok = tok; a.geo = geo; a.wblob = blob; a.ln1_w = w.ln1_w; a.ln1_b = w.ln1_b; a.bo = w.attn_out_b; a.ln_w = w.ln2_w;
  a.ln_b = w.ln2_b; a.b1 = w.ff1_b; a.b2 = w.ff2_b; a.xp = xp; a.vp = vp; a.dbg_tok = dbg_tok; a.p0 = p0; a.dbg_max = dbg_max; a.np = np; a.dc = dc;
  static bool attr_done = false;
  static int num_sms = 148;
  if (!attr_done) {
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_xformer_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xb::kSmemBytes));
    int dev = 0;
    SHERF_CUDA_OK(cudaGetDevice(&dev));
    SHERF_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_done = true;
  }
  const int ntiles = (np + 127) / 128;
  const int grid = ntiles < 2 * num_sms ? ntiles : 2 * num_sms;             // two co-resident CTAs per SM
  k_xformer_bf16<<<grid, 288, xb::kSmemBytes, st>>>(a);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

}  // namespace sherf
