// Fused feature fusion on the tensor cores: conv1d_projection 192 -> 96 (renderer.py:350), the per-token
// conv1d_reprojection 96 -> 32 over [tri_k | f2d_k | f3d_k] (renderer.py:423-424) and the transformer's first LayerNorm
// (renderer.py:931), for 128-point tiles, one persistent CTA per SM.  The gathered features are STREAMED from global
// memory in 32-column chunks through a double-buffered operand slot (hi part in shared memory, lo part in tensor
// memory); the projected 3-D feature never leaves the SM.  Outputs: tokens [3np][32] and LayerNorm-ed tokens [3np][32].
// Warp roles: warps 0-7 = chunk loaders + epilogues (row = TMEM lane, 16-column half per warp group),
//             warp 8 = converged MMA issuer, warp 9 lane 0 = TMA producer of the projection weights (3-stage ring).
#include "common.cuh"
#include "stages.cuh"
#include "umma.cuh"

namespace sherf {

namespace ff {
constexpr uint32_t kLbo = 2064;
constexpr uint32_t kCh0 = 0, kCh1 = 8 * kLbo;             // streamed operand chunk (8 core-matrix columns = 32 k) x 2
constexpr uint32_t kF3d = 16 * kLbo;                      // projected 3-D feature, 24 core-matrix columns
constexpr uint32_t kWr = 40 * kLbo;                       // resident reprojection weights [24 kg][32][4] hi | lo
constexpr uint32_t kWrBytes = 24 * 32 * 16;
constexpr uint32_t kRing = kWr + 2 * kWrBytes;            // projection weight ring
constexpr uint32_t kStage = 2 * 8 * 96 * 16;              // hi + lo of one 32-k chunk of Wp
constexpr int kNst = 3;
constexpr uint32_t kSmemBytes = kRing + kNst * kStage;
constexpr uint32_t kD1 = 0, kD2 = 96, kChLo0 = 192, kChLo1 = 224, kF3dLo = 256;     // tensor-memory columns
}  // namespace ff

struct FfArgs {
  const float *f3raw, *comb;           // [np][192], [np][288] (token k at k*96: tri_k(32) | f2d_k(32) | unused(32))
  const unsigned char* wblob;          // 6 projection chunks (hi|lo each), then reprojection hi | lo
  const float *bp, *br, *ln_w, *ln_b;  // conv1d_projection bias [96], conv1d_reprojection bias [32], LayerNorm-1 affine [32]
  float *tok, *ln;                     // [3np][32] each
  int np;
};

__device__ __forceinline__ void ff_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

template <int PREC>
__global__ void __launch_bounds__(320, 1) k_fusion_fused(const FfArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t wfull[ff::kNst], wempty[ff::kNst], chfull[2], chempty[2], acc1, acc2, f3d_ready;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bp[96], s_br[32], s_lnw[32], s_lnb[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < ff::kNst; ++s) { umma::mbar_init(&wfull[s], 1); umma::mbar_init(&wempty[s], 1); }
    for (int b = 0; b < 2; ++b) { umma::mbar_init(&chfull[b], 256); umma::mbar_init(&chempty[b], 1); }
    umma::mbar_init(&acc1, 1); umma::mbar_init(&acc2, 1); umma::mbar_init(&f3d_ready, 256);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  // resident reprojection weights (hi | lo) follow the six projection chunks in the blob
  for (int i = tid; i < (int)(2 * ff::kWrBytes / 16); i += blockDim.x)
    reinterpret_cast<float4*>(smem + ff::kWr)[i] = __ldg(reinterpret_cast<const float4*>(a.wblob + 6 * ff::kStage) + i);
  if (tid < 96) s_bp[tid] = a.bp[tid];
  if (tid < 32) { s_br[tid] = a.br[tid]; s_lnw[tid] = a.ln_w[tid]; s_lnb[tid] = a.ln_b[tid]; }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int ntiles = (a.np + 127) / 128;

  if (warp == 9) {
    // ===================== TMA producer: projection weight chunks =====================
    if (lane == 0) {
      uint32_t wc = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int c = 0; c < 6; ++c, ++wc) {
          const int s = wc % ff::kNst;
          umma::mbar_wait(&wempty[s], ((wc / ff::kNst) & 1) ^ 1);
          const uint32_t bytes = (PREC == 3) ? ff::kStage : ff::kStage / 2;
          umma::mbar_arrive_expect_tx(&wfull[s], bytes);
          umma::bulk_g2s(smem + ff::kRing + s * ff::kStage, a.wblob + (size_t)c * ff::kStage, bytes, &wfull[s]);
        }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer (converged warp) =====================
    const uint32_t sbase = umma::smem_u32(smem);
    const uint32_t el = umma::elect_one();         // one lane issues every MMA / commit of this warp
    uint32_t cnt = 0, wc = 0, par_t = 0;
    auto gemm = [&](uint32_t a_hi_addr, uint32_t a_lo_col, uint32_t w_hi_addr, uint32_t w_lo_addr, int N, uint32_t dcol, uint32_t acc0) {
      const uint32_t idesc = umma::make_idesc_tf32(128, N);
      const uint32_t w_lbo = (uint32_t)N * 16u;
      const uint64_t ah0 = umma::make_smem_desc(a_hi_addr, ff::kLbo, 128u);
      const uint64_t wh0 = umma::make_smem_desc(w_hi_addr, w_lbo, 128u);
      const uint64_t wl0 = umma::make_smem_desc(w_lo_addr, w_lbo, 128u);
      const uint64_t da = (uint64_t)((2u * ff::kLbo) >> 4), dw = (uint64_t)((2u * w_lbo) >> 4);
#pragma unroll
      for (int st = 0; st < 4; ++st) {                       // 8 core-matrix columns = 4 MMA k-steps
        const uint32_t acc = st == 0 ? acc0 : 1u;
        if (PREC == 3) {
          umma::mma_tf32_ts_e(tmem_base + dcol, tmem_base + a_lo_col + (uint32_t)st * 8u, wh0 + (uint64_t)st * dw, idesc, acc, el);
          umma::mma_tf32_ss_e(tmem_base + dcol, ah0 + (uint64_t)st * da, wl0 + (uint64_t)st * dw, idesc, 1u, el);
          umma::mma_tf32_ss_e(tmem_base + dcol, ah0 + (uint64_t)st * da, wh0 + (uint64_t)st * dw, idesc, 1u, el);
        } else {
          umma::mma_tf32_ss_e(tmem_base + dcol, ah0 + (uint64_t)st * da, wh0 + (uint64_t)st * dw, idesc, acc, el);
        }
      }
    };
    const uint32_t wr_hi = sbase + ff::kWr, wr_lo = wr_hi + ff::kWrBytes;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      // ---- projection: D1[128 x 96] = f3raw[128 x 192] * Wp^T, six streamed chunks ----
      for (int c = 0; c < 6; ++c, ++cnt, ++wc) {
        const int b = cnt & 1, s = wc % ff::kNst;
        umma::mbar_wait(&chfull[b], (cnt >> 1) & 1);
        umma::mbar_wait(&wfull[s], (wc / ff::kNst) & 1);
        umma::tc_fence_after_sync();
        const uint32_t w_hi = sbase + ff::kRing + (uint32_t)s * ff::kStage;
        gemm(sbase + (b ? ff::kCh1 : ff::kCh0), b ? ff::kChLo1 : ff::kChLo0, w_hi, w_hi + ff::kStage / 2, 96, ff::kD1, c == 0 ? 0u : 1u);
        umma::mma_commit_e(&chempty[b], el);
        umma::mma_commit_e(&wempty[s], el);
      }
      umma::mma_commit_e(&acc1, el);
      // ---- reprojection per token: D2[:, 32t:+32] = f3d_t * Wr[:,64:96]^T + tri_t * Wr[:,0:32]^T + f2d_t * Wr[:,32:64]^T ----
      umma::mbar_wait(&f3d_ready, par_t);
      umma::tc_fence_after_sync();
      for (int t = 0; t < 3; ++t) {
        gemm(sbase + ff::kF3d + (uint32_t)(8 * t) * ff::kLbo, ff::kF3dLo + (uint32_t)(32 * t), wr_hi + 16u * 32u * 16u, wr_lo + 16u * 32u * 16u, 32,
             ff::kD2 + (uint32_t)(32 * t), 0u);
        for (int part = 0; part < 2; ++part, ++cnt) {
          const int b = cnt & 1;
          umma::mbar_wait(&chfull[b], (cnt >> 1) & 1);
          umma::tc_fence_after_sync();
          gemm(sbase + (b ? ff::kCh1 : ff::kCh0), b ? ff::kChLo1 : ff::kChLo0, wr_hi + (uint32_t)(8 * part) * 32u * 16u,
               wr_lo + (uint32_t)(8 * part) * 32u * 16u, 32, ff::kD2 + (uint32_t)(32 * t), 1u);
          umma::mma_commit_e(&chempty[b], el);
        }
      }
      umma::mma_commit_e(&acc2, el);
      par_t ^= 1;
      __syncwarp();
    }
  } else {
    // ===================== chunk loaders + epilogues (warps 0-7) =====================
    const int q = warp & 3, hsel = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t tb = tmem_base + ((uint32_t)(32 * q) << 16);
    uint32_t cnt = 0, par_t = 0;
    auto split_store = [&](unsigned char* buf, int kg0, uint32_t lo_col, const float (&v)[16]) {
      uint32_t lo[16];
      float h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { h[i] = umma::to_tf32(v[i]); lo[i] = __float_as_uint(umma::to_tf32(v[i] - h[i])); }
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4)
        *reinterpret_cast<float4*>(buf + (kg0 + g4) * ff::kLbo + row * 16) = make_float4(h[4 * g4], h[4 * g4 + 1], h[4 * g4 + 2], h[4 * g4 + 3]);
      if (PREC == 3) umma::tmem_st16(tb + lo_col, lo);
    };
    // The gathered features are streamed in twelve 32-column chunks per tile (six of f3raw, then tri_t | f2d_t of the three tokens);
    // this thread handles 16 columns of its row.  The global loads run TWO chunks ahead of the operand-slot hand-off (software
    // pipeline across chunks and tiles, registers va / vb), so their latency overlaps the slot waits and the MMAs.
    auto issue_loads = [&](const float* src_row, bool ok, float (&v)[16]) {
      const float4* src = reinterpret_cast<const float4*>(src_row + 16 * hsel);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 f = ok ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
      }
    };
    auto commit_chunk = [&](const float (&v)[16]) {          // registers -> operand slot (cnt & 1)
      const int b = cnt & 1;
      umma::mbar_wait(&chempty[b], ((cnt >> 1) & 1) ^ 1);
      umma::tc_fence_after_sync();
      split_store(smem + (b ? ff::kCh1 : ff::kCh0), 4 * hsel, (b ? ff::kChLo1 : ff::kChLo0) + (uint32_t)(16 * hsel), v);
      if (PREC == 3) umma::tmem_st_wait();
      umma::fence_proxy_async_smem();
      umma::tc_fence_before_sync();
      ff_arrive(&chfull[b]);
      ++cnt;
    };
    auto chunk_src = [&](int m, int k) -> const float* {
      return k < 6 ? a.f3raw + (size_t)m * 192 + 32 * k : a.comb + (size_t)m * 288 + 96 * ((k - 6) >> 1) + 32 * ((k - 6) & 1);
    };
    float va[16], vb[16];
    {
      const int m0 = blockIdx.x * 128 + row;
      const bool ok0 = (int)blockIdx.x < ntiles && m0 < a.np;
      issue_loads(chunk_src(m0, 0), ok0, va);
      issue_loads(chunk_src(m0, 1), ok0, vb);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m = tile * 128 + row;
      const bool row_ok = m < a.np;
      const int mn = (tile + (int)gridDim.x) * 128 + row;                     // this thread's row in the CTA's next tile
      const bool next_ok = tile + (int)gridDim.x < ntiles && mn < a.np;
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        if (k == 6) {
          // ---- E1: projected 3-D feature = D1 + bias -> operand for the reprojection (48 columns per thread) ----
          umma::mbar_wait(&acc1, par_t);
          umma::tc_fence_after_sync();
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int c0 = 48 * hsel + 16 * i;
            uint32_t d[16];
            umma::tmem_ld16(tb + ff::kD1 + (uint32_t)c0, d);
            umma::tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(d[e]) + s_bp[c0 + e];
            split_store(smem + ff::kF3d, c0 / 4, ff::kF3dLo + (uint32_t)c0, v);
          }
          if (PREC == 3) umma::tmem_st_wait();
          umma::fence_proxy_async_smem();
          umma::tc_fence_before_sync();
          ff_arrive(&f3d_ready);
        }
        if (k & 1) {
          commit_chunk(vb);
          if (k + 2 < 12) issue_loads(chunk_src(m, k + 2), row_ok, vb); else issue_loads(chunk_src(mn, k + 2 - 12), next_ok, vb);
        } else {
          commit_chunk(va);
          if (k + 2 < 12) issue_loads(chunk_src(m, k + 2), row_ok, va); else issue_loads(chunk_src(mn, k + 2 - 12), next_ok, va);
        }
      }
      // ---- E2: tokens = D2 + bias -> global; LayerNorm-1 -> global.  hsel 0: tokens 0 and 2, hsel 1: token 1 ----
      umma::mbar_wait(&acc2, par_t);
      umma::tc_fence_after_sync();
      for (int t = hsel; t < 3; t += 2) {
        uint32_t d0[16], d1[16];
        umma::tmem_ld16(tb + ff::kD2 + (uint32_t)(32 * t), d0);
        umma::tmem_ld16(tb + ff::kD2 + (uint32_t)(32 * t + 16), d1);
        umma::tmem_ld_wait();
        float tk[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) { tk[e] = __uint_as_float(d0[e]) + s_br[e]; tk[16 + e] = __uint_as_float(d1[e]) + s_br[16 + e]; }
        float mean = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e) mean += tk[e];
        mean *= (1.f / 32.f);
        float var = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e) { const float dd = tk[e] - mean; var += dd * dd; }
        const float rstd = rsqrtf(var * (1.f / 32.f) + 1e-5f);
        if (row_ok) {
          float4* pt = reinterpret_cast<float4*>(a.tok + (size_t)(m * 3 + t) * 32);
          float4* pl = reinterpret_cast<float4*>(a.ln + (size_t)(m * 3 + t) * 32);
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            pt[g4] = make_float4(tk[4 * g4], tk[4 * g4 + 1], tk[4 * g4 + 2], tk[4 * g4 + 3]);
            pl[g4] = make_float4((tk[4 * g4] - mean) * rstd * s_lnw[4 * g4] + s_lnb[4 * g4],
                                 (tk[4 * g4 + 1] - mean) * rstd * s_lnw[4 * g4 + 1] + s_lnb[4 * g4 + 1],
                                 (tk[4 * g4 + 2] - mean) * rstd * s_lnw[4 * g4 + 2] + s_lnb[4 * g4 + 2],
                                 (tk[4 * g4 + 3] - mean) * rstd * s_lnw[4 * g4 + 3] + s_lnb[4 * g4 + 3]);
          }
        }
      }
      par_t ^= 1;
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// blob: 6 x [hi: 8 kg x 96 rows x 4 | lo] (projection chunks), then [hi: 24 kg x 32 rows x 4 | lo] (reprojection)
__global__ void k_pack_fusion(const float* __restrict__ wp, const float* __restrict__ wr, float* __restrict__ blob) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int kProjHalf = 8 * 96 * 4;                     // floats in one hi (or lo) part of a projection chunk
  constexpr int kProjAll = 6 * 2 * kProjHalf;
  constexpr int kReHalf = 24 * 32 * 4;
  if (i < 6 * kProjHalf) {
    const int c = i / kProjHalf, r = i % kProjHalf;
    const int e = r & 3, n = (r >> 2) % 96, kg = (r >> 2) / 96;
    const float v = wp[n * 192 + 32 * c + 4 * kg + e];
    const float h = umma::to_tf32(v);
    blob[c * 2 * kProjHalf + r] = h;
    blob[c * 2 * kProjHalf + kProjHalf + r] = umma::to_tf32(v - h);
  } else if (i < 6 * kProjHalf + kReHalf) {
    const int r = i - 6 * kProjHalf;
    const int e = r & 3, n = (r >> 2) % 32, kg = (r >> 2) / 32;
    const float v = wr[n * 96 + 4 * kg + e];
    const float h = umma::to_tf32(v);
    blob[kProjAll + r] = h;
    blob[kProjAll + kReHalf + r] = umma::to_tf32(v - h);
  }
}

size_t fusion_blob_floats() { return (size_t)6 * 2 * 8 * 96 * 4 + 2 * 24 * 32 * 4; }

int run_pack_fusion(const SherfWeights& w, float* blob, cudaStream_t st) {
  const int total = 6 * 8 * 96 * 4 + 24 * 32 * 4;
  k_pack_fusion<<<ceil_div(total, 256), 256, 0, st>>>(w.proj_w, w.reproj_w, blob);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

int run_fusion_fused(int prec, const SherfWeights& w, const float* blob, const float* f3raw, const float* comb, float* tok, float* ln,
                     int np, cudaStream_t st) {
  if (np <= 0) return SHERF_OK;
  FfArgs a;
  a.f3raw = f3raw; a.comb = comb; a.wblob = reinterpret_cast<const unsigned char*>(blob); a.bp = w.proj_b; a.br = w.reproj_b;
  a.ln_w = w.ln1_w; a.ln_b = w.ln1_b; a.tok = tok; a.ln = ln; a.np = np;
  static bool attr_done = false;
  static int num_sms = 148;
  if (!attr_done) {
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_fusion_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff::kSmemBytes));
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_fusion_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff::kSmemBytes));
    int dev = 0;
    SHERF_CUDA_OK(cudaGetDevice(&dev));
    SHERF_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_done = true;
  }
  const int ntiles = (np + 127) / 128;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  if (prec == 3) k_fusion_fused<3><<<grid, 320, ff::kSmemBytes, st>>>(a);
  else k_fusion_fused<1><<<grid, 320, ff::kSmemBytes, st>>>(a);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

}  // namespace sherf
