// Fused 3-token transformer layer on the tensor cores (renderer.py:920-993) + decoder-input assembly (renderer.py:432,
// triplane.py:293,308), for 128-POINT tiles, one persistent CTA per SM.  Rows of every MMA are points and the three
// tokens of a point live in different column blocks of the same TMEM lane, so the 3x3 attention, the LayerNorm, the
// residuals and the GELU are thread-local in the epilogues -- no cross-thread traffic at all.
//   phase            MMA (M=128 points)                         epilogue (thread = (point, query token t in {0,1}))
//   qkv, head h=0..2 D3[:, j*48:+48] = LN1_j * Wqkv_h^T  j=0..2   softmax(q_t k_j^T / 4) v_j -> ATT_t[:, h*16:+16]
//   to_out           D4[:, t*32:+32] = ATT_t * Wo^T               + bias + tok_t -> tok2_t ; LayerNorm -> LN2_t
//   ff1              D5[:, t*32:+32] = LN2_t * W1^T               GELU(. + b1) -> G_t
//   ff2              D6[:, t*32:+32] = G_t * W2^T                 + b2 + tok2_t -> tok3_t -> x / fv rows
// Token 2 is only a key/value source (the decoder reads tokens 0 and 1, triplane.py:288-289).
// Operands: hi parts in shared memory (UMMA K-major no-swizzle canonical layout, padded LBO), lo parts (3xTF32) in
// tensor memory; all weights (64 KB hi+lo) stay resident in shared memory for the lifetime of the CTA.
#include "common.cuh"
#include "stages.cuh"
#include "umma.cuh"

namespace sherf {

namespace xf {
constexpr uint32_t kLbo = 2064;
// shared memory map (bytes)
constexpr uint32_t kBuf1 = 0;                       // LN1 (3 tokens x 8 kg) -> LN2 (2 x 8 kg)
constexpr uint32_t kBuf2 = 24 * kLbo;               // ATT (2 tokens x 12 kg) -> G (2 x 8 kg)
constexpr uint32_t kW = 48 * kLbo;                  // resident weights, hi then lo per block
// weight blocks (floats, hi part; lo part follows at +kWFloats)
constexpr int kWqkv = 0;                            // 3 heads x [8 kg][48 rows][4]
constexpr int kWo = 3 * 8 * 48 * 4;                 // [12 kg][32][4]
constexpr int kW1 = kWo + 12 * 32 * 4;              // [8 kg][32][4]
constexpr int kW2 = kW1 + 8 * 32 * 4;
constexpr int kWFloats = kW2 + 8 * 32 * 4;          // 8192
constexpr uint32_t kSmemBytes = kW + 2 * kWFloats * 4;
// tensor memory map (columns)
constexpr uint32_t kD3 = 0, kD4 = 144, kD5 = 208, kLo1 = 272, kLo2 = 368;     // LN1_lo/LN2_lo at kLo1, ATT_lo/G_lo at kLo2
}  // namespace xf

struct XfArgs {
  const float *ln1, *tok, *geo;        // [3np][32] LayerNorm-ed tokens, [3np][32] tokens (residual), [np][8] can / cdir
  const float* pe;                     // [np][64] positional encodings made by k_point_pe: 36 of can (pos_enc) | 24 of cdir (view_enc) | 0 x 4
  const float* wblob;                  // [2][8192] canonical hi | lo
  const float *bo, *ln_w, *ln_b, *b1, *b2;
  float *x, *fv;                       // [np][72], [np][188] (columns 128..187)
  unsigned char *xp, *vp;              // optional packed bf16 hi/lo tiles for the ping-pong decoder (decoder_pp.cu); replace x / fv
  float* dbg_tok; int64_t p0, dbg_max;
  int np;
};

__device__ __forceinline__ void xf_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

template <int PREC>
__global__ void __launch_bounds__(288, 1) k_xformer_fused(const XfArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* buf1 = smem + xf::kBuf1;
  unsigned char* buf2 = smem + xf::kBuf2;
  float* w_hi = reinterpret_cast<float*>(smem + xf::kW);
  __shared__ __align__(8) uint64_t acc_bar, a_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bo[32], s_lnw[32], s_lnb[32], s_b1[32], s_b2[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) { umma::mbar_init(&acc_bar, 1); umma::mbar_init(&a_bar, 256); umma::fence_mbar_init(); }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < 2 * xf::kWFloats / 4; i += blockDim.x)
    reinterpret_cast<float4*>(w_hi)[i] = __ldg(reinterpret_cast<const float4*>(a.wblob) + i);
  if (tid < 32) { s_bo[tid] = a.bo[tid]; s_lnw[tid] = a.ln_w[tid]; s_lnb[tid] = a.ln_b[tid]; s_b1[tid] = a.b1[tid]; s_b2[tid] = a.b2[tid]; }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int ntiles = (a.np + 127) / 128;

  if (warp == 8) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    {
      uint32_t par_a = 0;
      const uint32_t el = umma::elect_one();       // one lane issues every MMA / commit of this warp
      const uint32_t b1s = umma::smem_u32(buf1), b2s = umma::smem_u32(buf2);
      const uint32_t whs = umma::smem_u32(w_hi), wls = whs + xf::kWFloats * 4;
      // one GEMM block: D[:, dcol:+N] (+)= A[128 x 4*nkg] * W[N x 4*nkg]^T
      auto gemm = [&](uint32_t a_hi_addr, uint32_t a_lo_col, int w_off_floats, int N, int nkg, uint32_t dcol) {
        const uint32_t idesc = umma::make_idesc_tf32(128, N);
        const uint32_t w_lbo = (uint32_t)N * 16u;
        const uint64_t ah0 = umma::make_smem_desc(a_hi_addr, xf::kLbo, 128u);
        const uint64_t wh0 = umma::make_smem_desc(whs + (uint32_t)w_off_floats * 4u, w_lbo, 128u);
        const uint64_t wl0 = umma::make_smem_desc(wls + (uint32_t)w_off_floats * 4u, w_lbo, 128u);
        const uint64_t da = (uint64_t)((2u * xf::kLbo) >> 4), dw = (uint64_t)((2u * w_lbo) >> 4);
        for (int st = 0; st < nkg / 2; ++st) {
          const uint32_t acc = st == 0 ? 0u : 1u;
          if (PREC == 3) {
            umma::mma_tf32_ts_e(tmem_base + dcol, tmem_base + a_lo_col + (uint32_t)st * 8u, wh0 + (uint64_t)st * dw, idesc, acc, el);
            umma::mma_tf32_ss_e(tmem_base + dcol, ah0 + (uint64_t)st * da, wl0 + (uint64_t)st * dw, idesc, 1u, el);
            umma::mma_tf32_ss_e(tmem_base + dcol, ah0 + (uint64_t)st * da, wh0 + (uint64_t)st * dw, idesc, 1u, el);
          } else {
            umma::mma_tf32_ss_e(tmem_base + dcol, ah0 + (uint64_t)st * da, wh0 + (uint64_t)st * dw, idesc, acc, el);
          }
        }
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int ph = 0; ph < 6; ++ph) {
          umma::mbar_wait(&a_bar, par_a);
          par_a ^= 1;
          umma::tc_fence_after_sync();
          if (ph < 3) {                       // qkv of head ph for the three tokens
            for (int j = 0; j < 3; ++j)
              gemm(b1s + (uint32_t)(j * 8) * xf::kLbo, xf::kLo1 + (uint32_t)(j * 32), xf::kWqkv + ph * 8 * 48 * 4, 48, 8, xf::kD3 + (uint32_t)(j * 48));
          } else if (ph == 3) {               // to_out for the two query tokens
            for (int t = 0; t < 2; ++t)
              gemm(b2s + (uint32_t)(t * 12) * xf::kLbo, xf::kLo2 + (uint32_t)(t * 48), xf::kWo, 32, 12, xf::kD4 + (uint32_t)(t * 32));
          } else if (ph == 4) {               // ff1
            for (int t = 0; t < 2; ++t)
              gemm(b1s + (uint32_t)(t * 8) * xf::kLbo, xf::kLo1 + (uint32_t)(t * 32), xf::kW1, 32, 8, xf::kD5 + (uint32_t)(t * 32));
          } else {                            // ff2
            for (int t = 0; t < 2; ++t)
              gemm(b2s + (uint32_t)(t * 8) * xf::kLbo, xf::kLo2 + (uint32_t)(t * 32), xf::kW2, 32, 8, xf::kD4 + (uint32_t)(t * 32));
          }
          umma::mma_commit_e(&acc_bar, el);
        }
      }
    }
  } else {
    // ===================== tile loader + epilogues (warps 0-7): thread = (point row, query token t) =====================
    const int q = warp & 3, t = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    const uint32_t tb = tmem_base + lane_base;
    uint32_t par_acc = 0;
    auto split_store = [&](unsigned char* buf, int kg0, uint32_t lo_col, const float (&v)[16]) {
      // 16 consecutive columns of this row: hi -> smem core-matrix columns kg0..kg0+3, lo -> TMEM
      uint32_t lo[16];
      float h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { h[i] = umma::to_tf32(v[i]); lo[i] = __float_as_uint(umma::to_tf32(v[i] - h[i])); }
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4)
        *reinterpret_cast<float4*>(buf + (kg0 + g4) * xf::kLbo + row * 16) = make_float4(h[4 * g4], h[4 * g4 + 1], h[4 * g4 + 2], h[4 * g4 + 3]);
      if (PREC == 3) umma::tmem_st16(tb + lo_col, lo);
    };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m = tile * 128 + row;
      const bool row_ok = m < a.np;
      // ---- LN1 tile: token t fully by this thread, token 2 split between the two threads of the row ----
      {
        float v[16];
        const float4* src = reinterpret_cast<const float4*>(a.ln1 + (size_t)(m * 3 + t) * 32);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = row_ok ? __ldg(src + half * 4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
          }
          split_store(buf1, t * 8 + half * 4, xf::kLo1 + (uint32_t)(t * 32 + half * 16), v);
        }
        const float4* src2 = reinterpret_cast<const float4*>(a.ln1 + (size_t)(m * 3 + 2) * 32) + t * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 f = row_ok ? __ldg(src2 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
        }
        split_store(buf1, 16 + t * 4, xf::kLo1 + (uint32_t)(64 + t * 16), v);
        if (PREC == 3) umma::tmem_st_wait();
      }
      umma::fence_proxy_async_smem();
      umma::tc_fence_before_sync();
      xf_arrive(&a_bar);
      // pull what this thread reads later into L2 while the tensor core works: the residual rows of this tile (to_out epilogue),
      // and the LayerNorm-ed rows / geometry of the CTA's next tile (the fusion kernel's output is larger than L2)
      {
        auto pf = [](const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
        if (row_ok) pf(a.tok + (size_t)(m * 3 + t) * 32);
        const int mn = m + (int)gridDim.x * 128;
        if (mn < a.np) {
          pf(a.ln1 + (size_t)(mn * 3 + t) * 32);
          if (t == 0) { pf(a.ln1 + (size_t)(mn * 3 + 2) * 32); pf(a.geo + (size_t)mn * 8); }
          pf(a.pe + (size_t)mn * 64 + 32 * t);
        }
      }

      // ---- three heads: attention of query token t over the three tokens ----
      for (int h = 0; h < 3; ++h) {
        umma::mbar_wait(&acc_bar, par_acc);
        par_acc ^= 1;
        umma::tc_fence_after_sync();
        uint32_t qv[16], kv[3][16], vv[3][16];
        umma::tmem_ld16(tb + xf::kD3 + (uint32_t)(t * 48), qv);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          umma::tmem_ld16(tb + xf::kD3 + (uint32_t)(j * 48 + 16), kv[j]);
          umma::tmem_ld16(tb + xf::kD3 + (uint32_t)(j * 48 + 32), vv[j]);
        }
        umma::tmem_ld_wait();
        float dots[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float s = 0.f;
#pragma unroll
          for (int d = 0; d < 16; ++d) s += __uint_as_float(qv[d]) * __uint_as_float(kv[j][d]);
          dots[j] = s * 0.25f;                              // dim_head ** -0.5   (renderer.py:956,971)
        }
        const float mx = fmaxf(dots[0], fmaxf(dots[1], dots[2]));
        const float e0 = expf(dots[0] - mx), e1 = expf(dots[1] - mx), e2 = expf(dots[2] - mx);
        const float inv = 1.f / (e0 + e1 + e2);
        const float a0 = e0 * inv, a1 = e1 * inv, a2 = e2 * inv;
        float att[16];
#pragma unroll
        for (int d = 0; d < 16; ++d)
          att[d] = a0 * __uint_as_float(vv[0][d]) + a1 * __uint_as_float(vv[1][d]) + a2 * __uint_as_float(vv[2][d]);
        split_store(buf2, t * 12 + h * 4, xf::kLo2 + (uint32_t)(t * 48 + h * 16), att);
        if (PREC == 3) umma::tmem_st_wait();
        umma::fence_proxy_async_smem();
        umma::tc_fence_before_sync();
        xf_arrive(&a_bar);
      }
      // ---- to_out + residual + LayerNorm ----
      float tok2[32];
      {
        umma::mbar_wait(&acc_bar, par_acc);
        par_acc ^= 1;
        umma::tc_fence_after_sync();
        uint32_t o0[16], o1[16];
        umma::tmem_ld16(tb + xf::kD4 + (uint32_t)(t * 32), o0);
        umma::tmem_ld16(tb + xf::kD4 + (uint32_t)(t * 32 + 16), o1);
        const float4* tr = reinterpret_cast<const float4*>(a.tok + (size_t)(m * 3 + t) * 32);
        float4 tk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) tk[i] = row_ok ? __ldg(tr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) { tok2[i] = __uint_as_float(o0[i]) + s_bo[i]; tok2[16 + i] = __uint_as_float(o1[i]) + s_bo[16 + i]; }
#pragma unroll
        for (int i = 0; i < 8; ++i) { tok2[4 * i] += tk[i].x; tok2[4 * i + 1] += tk[i].y; tok2[4 * i + 2] += tk[i].z; tok2[4 * i + 3] += tk[i].w; }
        float mean = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) mean += tok2[i];
        mean *= (1.f / 32.f);
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float d = tok2[i] - mean; var += d * d; }
        const float rstd = rsqrtf(var * (1.f / 32.f) + 1e-5f);
        float v[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (tok2[half * 16 + i] - mean) * rstd * s_lnw[half * 16 + i] + s_lnb[half * 16 + i];
          split_store(buf1, t * 8 + half * 4, xf::kLo1 + (uint32_t)(t * 32 + half * 16), v);
        }
        if (PREC == 3) umma::tmem_st_wait();
        umma::fence_proxy_async_smem();
        umma::tc_fence_before_sync();
        xf_arrive(&a_bar);
      }
      // ---- ff1 + GELU ----
      {
        umma::mbar_wait(&acc_bar, par_acc);
        par_acc ^= 1;
        umma::tc_fence_after_sync();
        uint32_t o0[16], o1[16];
        umma::tmem_ld16(tb + xf::kD5 + (uint32_t)(t * 32), o0);
        umma::tmem_ld16(tb + xf::kD5 + (uint32_t)(t * 32 + 16), o1);
        umma::tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float x = __uint_as_float(o0[i]) + s_b1[i]; v[i] = 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
        split_store(buf2, t * 8, xf::kLo2 + (uint32_t)(t * 32), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float x = __uint_as_float(o1[i]) + s_b1[16 + i]; v[i] = 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
        split_store(buf2, t * 8 + 4, xf::kLo2 + (uint32_t)(t * 32 + 16), v);
        if (PREC == 3) umma::tmem_st_wait();
        umma::fence_proxy_async_smem();
        umma::tc_fence_before_sync();
        xf_arrive(&a_bar);
      }
      // ---- ff2 + residual -> decoder inputs ----
      {
        // positional encodings of this row (renderer.py:432 pos_enc / view_enc), computed by k_point_pe at full occupancy: sixty sinf per
        // point inside this epilogue were a third of the kernel's instructions on its eight serialised warps (profiles/r1_q)
        float4 pev[9];
        {
          const float4* pr = reinterpret_cast<const float4*>(a.pe + (size_t)m * 64 + (t == 0 ? 0 : 36));
#pragma unroll
          for (int i = 0; i < 9; ++i) pev[i] = (row_ok && (t == 0 || i < 6)) ? __ldg(pr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        umma::mbar_wait(&acc_bar, par_acc);
        par_acc ^= 1;
        umma::tc_fence_after_sync();
        uint32_t o0[16], o1[16];
        umma::tmem_ld16(tb + xf::kD4 + (uint32_t)(t * 32), o0);
        umma::tmem_ld16(tb + xf::kD4 + (uint32_t)(t * 32 + 16), o1);
        umma::tmem_ld_wait();
        float tok3[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) { tok3[i] = __uint_as_float(o0[i]) + s_b2[i] + tok2[i]; tok3[16 + i] = __uint_as_float(o1[i]) + s_b2[16 + i] + tok2[16 + i]; }
        {
          const float g0 = row_ok ? a.geo[(size_t)m * 8 + 3 * t] : 0.f, g1 = row_ok ? a.geo[(size_t)m * 8 + 3 * t + 1] : 0.f,
                      g2 = row_ok ? a.geo[(size_t)m * 8 + 3 * t + 2] : 0.f;
          const float* pef = reinterpret_cast<const float*>(pev);              // pef[3 * mm + c] = sin(phase_mm + g_c * 2^(mm >> 1))
          // packed tile store: 8 consecutive k-columns of this row -> one 16-byte hi and one 16-byte lo chunk (consecutive rows
          // are consecutive chunks, so a warp writes 512 contiguous bytes per core-matrix column)
          auto put8 = [&](unsigned char* tile_base, int kg, uint32_t lo_off, const float* v8) {
            uint4 h, l;
            umma::split_bf16x2(v8[0], v8[1], h.x, l.x); umma::split_bf16x2(v8[2], v8[3], h.y, l.y);
            umma::split_bf16x2(v8[4], v8[5], h.z, l.z); umma::split_bf16x2(v8[6], v8[7], h.w, l.w);
            *reinterpret_cast<uint4*>(tile_base + (size_t)(kg * 128 + row) * 16) = h;
            *reinterpret_cast<uint4*>(tile_base + lo_off + (size_t)(kg * 128 + row) * 16) = l;
          };
          if (t == 0) {
            float vals[80];
            vals[0] = g0; vals[1] = g1; vals[2] = g2;
#pragma unroll
            for (int o = 0; o < 36; ++o) vals[3 + o] = pef[o];
#pragma unroll
            for (int o = 0; o < 32; ++o) vals[39 + o] = tok3[o];
#pragma unroll
            for (int o = 71; o < 80; ++o) vals[o] = 0.f;
            if (a.xp) {
              if (!row_ok) {
#pragma unroll
                for (int o = 0; o < 71; ++o) vals[o] = 0.f;
              }
              unsigned char* tb_ = a.xp + (size_t)tile * 40960;
#pragma unroll
              for (int kg = 0; kg < 10; ++kg) put8(tb_, kg, 20480u, vals + 8 * kg);
            } else if (row_ok) {
              float4* dx = reinterpret_cast<float4*>(a.x + (size_t)m * 72);
#pragma unroll
              for (int i = 0; i < 18; ++i) dx[i] = make_float4(vals[4 * i], vals[4 * i + 1], vals[4 * i + 2], vals[4 * i + 3]);
            }
          } else {
            float vals[64];
            vals[0] = g0; vals[1] = g1; vals[2] = g2;
#pragma unroll
            for (int o = 0; o < 24; ++o) vals[3 + o] = pef[o];
#pragma unroll
            for (int o = 0; o < 32; ++o) vals[27 + o] = tok3[o];
#pragma unroll
            for (int o = 59; o < 64; ++o) vals[o] = 0.f;
            if (a.vp) {
              if (!row_ok) {
#pragma unroll
                for (int o = 0; o < 59; ++o) vals[o] = 0.f;
              }
              unsigned char* tb_ = a.vp + (size_t)tile * 32768;
#pragma unroll
              for (int kg = 0; kg < 8; ++kg) put8(tb_, kg, 16384u, vals + 8 * kg);
            } else if (row_ok) {
              float4* dv = reinterpret_cast<float4*>(a.fv + (size_t)m * 188 + 128);
#pragma unroll
              for (int i = 0; i < 15; ++i) dv[i] = make_float4(vals[4 * i], vals[4 * i + 1], vals[4 * i + 2], vals[4 * i + 3]);
            }
          }
        }
        if (row_ok) {
          if (a.dbg_tok && a.p0 + m < a.dbg_max) {
#pragma unroll
            for (int o = 0; o < 32; ++o) a.dbg_tok[(a.p0 + m) * 64 + t * 32 + o] = tok3[o];
          }
        }
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// Positional encodings of the canonical position (pos_enc, 6 octaves) and view direction (view_enc, 4 octaves) of every point
// (renderer.py:432, PositionalEncoding :900-916): pe[p][3*mm + c] = sin(phase(mm) + x_c * 2^(mm >> 1)), phase = 0 | pi/2 -- the same
// separately rounded multiply and add as torch.addcmul.  One thread per value: the sixty sinf per point run at full occupancy here.
__global__ void __launch_bounds__(256) k_point_pe(const float* __restrict__ geo, float* __restrict__ pe, int np_host, const DevCount dc) {
  const int np = resolve_np(np_host, dc);
  // thread = (point, slot) with 32 slots per point: slot s < 18 -> position coordinate c = s % 3 at octave k = s / 3 (6 octaves),
  // 18 <= s < 30 -> direction coordinate at octave (s - 18) / 3 (4 octaves); each thread writes the phase-0 and the phase-pi/2 value
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)np * 32) return;
  const int p = (int)(idx >> 5), s = (int)(idx & 31);
  float* row = pe + (size_t)p * 64;
  if (s >= 30) { row[60 + 2 * (s - 30)] = 0.f; row[61 + 2 * (s - 30)] = 0.f; return; }
  const bool dir = s >= 18;
  const int ss = dir ? s - 18 : s;
  const int k = (ss * 11) >> 5, c = ss - 3 * k;                    // ss / 3 for ss < 18
  const float g = geo[(size_t)p * 8 + (dir ? 3 : 0) + c];
  const float y = __fmul_rn(g, (float)(1 << k));
  float* o = row + (dir ? 36 : 0) + 6 * k + c;                     // pe[3 * mm + c], mm = 2k (phase 0) and 2k + 1 (phase pi/2)
  o[0] = sinf(__fadd_rn(0.f, y));                                  // torch.addcmul(0, x, f): 0 + y (turns -0 into +0)
  o[3] = sinf(__fadd_rn(kPi2, y));
}

int run_point_pe(const float* geo, float* pe, int np, cudaStream_t st, DevCount dc) {
  if (np <= 0) return SHERF_OK;
  k_point_pe<<<ceil_div((int64_t)np * 32, 256), 256, 0, st>>>(geo, pe, np, dc);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// canonical weight blob: element (kg, n, e) of a block = W_block[n][kg*4 + e]
__global__ void k_pack_xformer(const float* __restrict__ wqkv, const float* __restrict__ wo, const float* __restrict__ w1,
                               const float* __restrict__ w2, float* __restrict__ blob) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= xf::kWFloats) return;
  float v;
  if (i < xf::kWo) {                                        // head h block: rows [q_h(16) | k_h(16) | v_h(16)] of to_qkv.weight [144][32]
    const int h = i / (8 * 48 * 4), r = i % (8 * 48 * 4);
    const int e = r & 3, n = (r >> 2) % 48, kg = (r >> 2) / 48;
    const int src_row = (n / 16) * 48 + h * 16 + (n % 16);  // q block at 0, k at 48, v at 96; head h at +16h   (renderer.py:968-969)
    v = wqkv[src_row * 32 + kg * 4 + e];
  } else if (i < xf::kW1) {
    const int r = i - xf::kWo;
    const int e = r & 3, n = (r >> 2) % 32, kg = (r >> 2) / 32;
    v = wo[n * 48 + kg * 4 + e];
  } else if (i < xf::kW2) {
    const int r = i - xf::kW1;
    const int e = r & 3, n = (r >> 2) % 32, kg = (r >> 2) / 32;
    v = w1[n * 32 + kg * 4 + e];
  } else {
    const int r = i - xf::kW2;
    const int e = r & 3, n = (r >> 2) % 32, kg = (r >> 2) / 32;
    v = w2[n * 32 + kg * 4 + e];
  }
  const float h = umma::to_tf32(v);
  blob[i] = h;
  blob[xf::kWFloats + i] = umma::to_tf32(v - h);
}

size_t xformer_blob_floats() { return 2 * (size_t)xf::kWFloats; }

int run_pack_xformer(const SherfWeights& w, float* blob, cudaStream_t st) {
  k_pack_xformer<<<ceil_div(xf::kWFloats, 256), 256, 0, st>>>(w.qkv_w, w.attn_out_w, w.ff1_w, w.ff2_w, blob);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

int run_xformer_fused(int prec, const SherfWeights& w, const float* blob, const float* ln1, const float* tok, const float* geo, float* x,
                      float* fv, int np, float* dbg_tok, int64_t p0, int64_t dbg_max, cudaStream_t st, unsigned char* xp, unsigned char* vp,
                      float* pe_buf) {
  if (np <= 0) return SHERF_OK;
  { const int rc = run_point_pe(geo, pe_buf, np, st); if (rc) return rc; }
  XfArgs a;
  a.pe = pe_buf;
  a.ln1 = ln1; a.tok = tok; a.geo = geo; a.wblob = blob; a.bo = w.attn_out_b; a.ln_w = w.ln2_w; a.ln_b = w.ln2_b; a.b1 = w.ff1_b;
  a.b2 = w.ff2_b; a.x = x; a.fv = fv; a.xp = xp; a.vp = vp; a.dbg_tok = dbg_tok; a.p0 = p0; a.dbg_max = dbg_max; a.np = np;
  static bool attr_done = false;
  static int num_sms = 148;
  if (!attr_done) {
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_xformer_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xf::kSmemBytes));
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_xformer_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xf::kSmemBytes));
    int dev = 0;
    SHERF_CUDA_OK(cudaGetDevice(&dev));
    SHERF_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_done = true;
  }
  const int ntiles = (np + 127) / 128;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  if (prec == 3) k_xformer_fused<3><<<grid, 288, xf::kSmemBytes, st>>>(a);
  else k_xformer_fused<1><<<grid, 288, xf::kSmemBytes, st>>>(a);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

}  // namespace sherf
