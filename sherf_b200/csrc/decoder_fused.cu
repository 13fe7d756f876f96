// Fused NeRF-decoder trunk on the tensor cores: pts_linears[0..7] (with the skip concat) + feature_linear + alpha_linear
// (triplane.py:293-303) for 128-point tiles, one persistent CTA per SM.  Activations never leave the SM:
//   hi parts   : shared memory, UMMA K-major no-swizzle canonical layout (padded LBO, see mlp_umma.cu)
//   lo parts   : tensor memory (3xTF32 error compensation), consumed by tcgen05.mma with the A operand in TMEM
//   accumulator: tensor memory, read back by the epilogue warps (bias + ReLU + tf32 split) straight into the next layer's operands
// Weights stream from L2 through a 3-stage shared-memory ring filled by a TMA (cp.async.bulk) producer thread.
// Warp roles: warps 0-7 = tile loader + epilogue (warp w owns TMEM lane quarter w & 3, column half w >> 2),
//             warp 8 lane 0 = MMA issuer, warp 9 lane 0 = TMA weight producer.
#include "common.cuh"
#include "stages.cuh"
#include "umma.cuh"

namespace sherf {

constexpr int kFusedChunks = 38;
constexpr int kFusedLayers = 9;
constexpr int kNst = 3;                                  // weight ring stages
constexpr uint32_t kLbo = 2064;                          // padded K-direction stride of A operands (bytes)
constexpr uint32_t kXBytes = 18 * kLbo, kHBytes = 32 * kLbo;
constexpr uint32_t kStageBytes = 2 * 8 * 144 * 16;       // hi + lo, 8 core-matrix columns, up to 144 rows
constexpr uint32_t kColD = 0, kColHlo = 160, kColXlo = 288;

static_assert(kFusedChunks == 38 && kFusedLayers == 9, "FusedSchedule (stages.cuh) is sized for 38 chunks / 9 layers");

struct FusedArgs {
  const float* X; int ldx;             // [np][72] decoder input rows (PE6(can) | tok0 | 0)
  const unsigned char* wblob;          // packed chunks in schedule order
  const float* bias;                   // [9][144] (row 8: feature bias 0..127, alpha bias at 128)
  float* fv; int ldfv;                 // feature output -> fv[:, 0:128]
  float* sigma;                        // [np]
  int np;
  FusedSchedule sch;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

template <int PREC>
__global__ void __launch_bounds__(320, 1) k_decoder_fused(const FusedArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* X_hi = smem;
  unsigned char* H_hi = smem + kXBytes;
  unsigned char* Wst = smem + kXBytes + kHBytes;
  float* s_bias = reinterpret_cast<float*>(Wst + kNst * kStageBytes);
  __shared__ __align__(8) uint64_t full_bar[kNst], empty_bar[kNst], acc_bar, a_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kNst; ++s) { umma::mbar_init(&full_bar[s], 1); umma::mbar_init(&empty_bar[s], 1); }
    umma::mbar_init(&acc_bar, 1);
    umma::mbar_init(&a_bar, 256);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < kFusedLayers * 144; i += blockDim.x) s_bias[i] = a.bias[i];
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int ntiles = (a.np + 127) / 128;

  if (warp == 9) {
    // ===================== TMA weight producer =====================
    if (lane == 0) {
      uint32_t cc = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int c = 0; c < kFusedChunks; ++c, ++cc) {
          const int s = cc % kNst;
          umma::mbar_wait(&empty_bar[s], ((cc / kNst) & 1) ^ 1);
          const FusedChunk& ch = a.sch.ch[c];
          const uint32_t bytes = (PREC == 3) ? ch.w_bytes : ch.w_bytes / 2;      // single-pass TF32 needs the hi half only
          umma::mbar_arrive_expect_tx(&full_bar[s], bytes);
          umma::bulk_g2s(Wst + s * kStageBytes, a.wblob + ch.w_off, bytes, &full_bar[s]);
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t cc = 0, par_a = 0;
      const uint32_t x_s = umma::smem_u32(X_hi), h_s = umma::smem_u32(H_hi), w_s = umma::smem_u32(Wst);
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int c = 0; c < kFusedChunks; ++c, ++cc) {
          const FusedChunk& ch = a.sch.ch[c];
          if (ch.first) { umma::mbar_wait(&a_bar, par_a); par_a ^= 1; umma::tc_fence_after_sync(); }
          const int s = cc % kNst;
          umma::mbar_wait(&full_bar[s], (cc / kNst) & 1);
          umma::tc_fence_after_sync();
          const uint32_t Np = a.sch.layer_np[ch.layer];
          const uint32_t idesc = umma::make_idesc_tf32(128, (int)Np);
          const uint32_t a_base = (ch.src == 0 ? x_s : h_s) + (uint32_t)ch.kg0 * kLbo;
          const uint32_t alo_col = tmem_base + (ch.src == 0 ? kColXlo : kColHlo) + (uint32_t)ch.kg0 * 4u;
          const uint32_t w_hi = w_s + (uint32_t)s * kStageBytes, w_lo = w_hi + (uint32_t)ch.nkg * Np * 16u;
          const uint32_t w_lbo = Np * 16u;
          for (int st = 0; st < ch.nkg / 2; ++st) {
            const uint64_t ah = umma::make_smem_desc(a_base + (uint32_t)st * 2u * kLbo, kLbo, 128u);
            const uint64_t wh = umma::make_smem_desc(w_hi + (uint32_t)st * 2u * w_lbo, w_lbo, 128u);
            const uint32_t acc = (ch.first && st == 0) ? 0u : 1u;
            if (PREC == 3) {
              const uint64_t wl = umma::make_smem_desc(w_lo + (uint32_t)st * 2u * w_lbo, w_lbo, 128u);
              umma::mma_tf32_ts(tmem_base + kColD, alo_col + (uint32_t)st * 8u, wh, idesc, acc);
              umma::mma_tf32_ss(tmem_base + kColD, ah, wl, idesc, 1u);
              umma::mma_tf32_ss(tmem_base + kColD, ah, wh, idesc, 1u);
            } else {
              umma::mma_tf32_ss(tmem_base + kColD, ah, wh, idesc, acc);
            }
          }
          umma::mma_commit(&empty_bar[s]);                 // weight stage reusable once these MMAs retire
          if (ch.last) umma::mma_commit(&acc_bar);         // layer accumulator complete
        }
      }
    }
  } else {
    // ===================== tile loader + epilogue (warps 0-7) =====================
    const int q = warp & 3, hsel = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    uint32_t par_acc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m = tile * 128 + row;
      const bool row_ok = m < a.np;
      // ---- X tile: 72 columns = 18 core-matrix columns; this thread handles 9 of them for its row ----
      {
        const float* xr = a.X + (size_t)m * a.ldx;
        for (int kg = hsel * 9; kg < hsel * 9 + 9; ++kg) {
          float4 v = row_ok ? *reinterpret_cast<const float4*>(xr + kg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 h = make_float4(umma::to_tf32(v.x), umma::to_tf32(v.y), umma::to_tf32(v.z), umma::to_tf32(v.w));
          *reinterpret_cast<float4*>(X_hi + kg * kLbo + row * 16) = h;
        }
        if (PREC == 3) {
          // lo part, 8 columns per tcgen05.st: thread covers columns [hsel*36, hsel*36+36) -> 4 full groups + a half group;
          // simpler and exact: hsel 0 writes groups 0..4 (cols 0..39), hsel 1 writes groups 5..8 (cols 40..71)
          const int g0 = hsel == 0 ? 0 : 5, g1 = hsel == 0 ? 5 : 9;
          for (int gi = g0; gi < g1; ++gi) {
            uint32_t lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float x = row_ok ? xr[gi * 8 + e] : 0.f;
              lo[e] = __float_as_uint(umma::to_tf32(x - umma::to_tf32(x)));
            }
            umma::tmem_st8(tmem_base + lane_base + kColXlo + (uint32_t)(gi * 8), lo);
          }
          umma::tmem_st_wait();
        }
      }
      umma::fence_proxy_async_smem();
      umma::tc_fence_before_sync();
      mbar_arrive(&a_bar);

      for (int l = 0; l < kFusedLayers; ++l) {
        umma::mbar_wait(&acc_bar, par_acc);
        par_acc ^= 1;
        umma::tc_fence_after_sync();
        const int Np = a.sch.layer_np[l];
        const int half = Np / 2;
        const float* bl = s_bias + l * 144;
        for (int j = 0; j < half / 8; ++j) {
          const int c0 = hsel * half + 8 * j;
          uint32_t v[8];
          umma::tmem_ld8(tmem_base + lane_base + kColD + (uint32_t)c0, v);
          umma::tmem_ld_wait();
          if (l < kFusedLayers - 1) {
            float x[8];
            uint32_t lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              x[i] = fmaxf(__uint_as_float(v[i]) + bl[c0 + i], 0.f);
              const float h = umma::to_tf32(x[i]);
              lo[i] = __float_as_uint(umma::to_tf32(x[i] - h));
              x[i] = h;
            }
            *reinterpret_cast<float4*>(H_hi + (c0 / 4) * kLbo + row * 16) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4*>(H_hi + (c0 / 4 + 1) * kLbo + row * 16) = make_float4(x[4], x[5], x[6], x[7]);
            if (PREC == 3) umma::tmem_st8(tmem_base + lane_base + kColHlo + (uint32_t)c0, lo);
          } else if (row_ok) {
            // feature_linear (cols 0..127, no activation) -> fv ; alpha_linear (col 128) -> sigma      triplane.py:302-303
            if (c0 < 128) {
              float4* dst = reinterpret_cast<float4*>(a.fv + (size_t)m * a.ldfv + c0);
              dst[0] = make_float4(__uint_as_float(v[0]) + bl[c0], __uint_as_float(v[1]) + bl[c0 + 1], __uint_as_float(v[2]) + bl[c0 + 2],
                                   __uint_as_float(v[3]) + bl[c0 + 3]);
              dst[1] = make_float4(__uint_as_float(v[4]) + bl[c0 + 4], __uint_as_float(v[5]) + bl[c0 + 5], __uint_as_float(v[6]) + bl[c0 + 6],
                                   __uint_as_float(v[7]) + bl[c0 + 7]);
            } else if (c0 == 128) {
              a.sigma[m] = __uint_as_float(v[0]) + bl[128];
            }
          }
        }
        if (l < kFusedLayers - 1) {
          if (PREC == 3) umma::tmem_st_wait();
          umma::fence_proxy_async_smem();
          umma::tc_fence_before_sync();
          mbar_arrive(&a_bar);
        }
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight packing in schedule order.  Chunk blob = [hi: nkg x Np x 4 floats][lo: same]; element (kg, n, e) = W[n][col0 + kg*4 + e].
struct FusedPackJob { const float* W; const float* Wextra; int N, ldw, col0, ncols, nkg, Np; uint32_t off; };
struct FusedPackJobs { FusedPackJob j[kFusedChunks]; };

__global__ void k_pack_fused(const FusedPackJobs jobs, unsigned char* blob) {
  const FusedPackJob jb = jobs.j[blockIdx.y];
  float* hi = reinterpret_cast<float*>(blob + jb.off);
  float* lo = hi + jb.nkg * jb.Np * 4;
  const int total = jb.nkg * jb.Np * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 3, n = (i >> 2) % jb.Np, kg = (i >> 2) / jb.Np;
    const int kk = kg * 4 + e;
    float v = 0.f;
    if (kk < jb.ncols) {
      if (n < jb.N) v = jb.W[(size_t)n * jb.ldw + jb.col0 + kk];
      else if (n == jb.N && jb.Wextra) v = jb.Wextra[jb.col0 + kk];
    }
    const float h = umma::to_tf32(v);
    hi[i] = h;
    lo[i] = umma::to_tf32(v - h);
  }
}

__global__ void k_fused_bias(const SherfWeights w, float* bias) {
  const int l = blockIdx.x, n = threadIdx.x;     // 9 x 144
  float v = 0.f;
  if (l < 8) { if (n < 128) v = w.pts_b[l][n]; }
  else { if (n < 128) v = w.feature_b[n]; else if (n == 128) v = w.alpha_b[0]; }
  bias[l * 144 + n] = v;
}

size_t fused_blob_bytes() { return (size_t)2 * (128 * (72 + 128 * 4 + 200 + 128 * 2) + 144 * 128) * 4 + 1024; }

int run_pack_fused(const SherfWeights& w, unsigned char* blob, float* bias, FusedSchedule& sch, cudaStream_t st) {
  FusedPackJobs jobs;
  int c = 0;
  uint32_t off = 0;
  auto seg = [&](int layer, const float* W, const float* Wextra, int N, int Np, int ldw, int col0, int ncols, int src, bool first_seg,
                 bool last_seg) {
    const int padded = (ncols + 7) / 8 * 8;                       // whole MMA k-steps
    for (int k0 = 0; k0 < padded; k0 += 32) {
      const int nkg = (padded - k0 >= 32 ? 32 : padded - k0) / 4;
      FusedChunk& ch = sch.ch[c];
      ch.src = (uint16_t)src; ch.kg0 = (uint16_t)(k0 / 4); ch.nkg = (uint16_t)nkg; ch.layer = (uint16_t)layer;
      ch.first = (uint16_t)(first_seg && k0 == 0); ch.last = (uint16_t)(last_seg && k0 + 32 >= padded);
      ch.w_off = off; ch.w_bytes = (uint32_t)(2 * nkg * Np * 16);
      FusedPackJob& j = jobs.j[c];
      j.W = W; j.Wextra = Wextra; j.N = N; j.ldw = ldw; j.col0 = col0 + k0; j.ncols = ncols - k0; j.nkg = nkg; j.Np = Np; j.off = off;
      off += ch.w_bytes;
      ++c;
    }
  };
  seg(0, w.pts_w[0], nullptr, 128, 128, 71, 0, 71, 0, true, true);
  for (int l = 1; l <= 4; ++l) seg(l, w.pts_w[l], nullptr, 128, 128, 128, 0, 128, 1, true, true);
  seg(5, w.pts_w[5], nullptr, 128, 128, 199, 0, 71, 0, true, false);       // skip concat: h = cat([x, h])   triplane.py:299-300
  seg(5, w.pts_w[5], nullptr, 128, 128, 199, 71, 128, 1, false, true);
  seg(6, w.pts_w[6], nullptr, 128, 128, 128, 0, 128, 1, true, true);
  seg(7, w.pts_w[7], nullptr, 128, 128, 128, 0, 128, 1, true, true);
  seg(8, w.feature_w, w.alpha_w, 128, 144, 128, 0, 128, 1, true, true);
  if (c != kFusedChunks) { set_error("internal: fused schedule has %d chunks", c); return SHERF_E_INVALID; }
  for (int l = 0; l < 8; ++l) sch.layer_np[l] = 128;
  sch.layer_np[8] = 144;
  sch.pad = 0;
  k_pack_fused<<<dim3(8, kFusedChunks), 256, 0, st>>>(jobs, blob);
  SHERF_LAUNCH_CHECK();
  k_fused_bias<<<kFusedLayers, 144, 0, st>>>(w, bias);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

int run_decoder_fused(int prec, const FusedSchedule& sch, const unsigned char* blob, const float* bias, const float* X, int ldx, float* fv,
                      int ldfv, float* sigma, int np, cudaStream_t st) {
  if (np <= 0) return SHERF_OK;
  FusedArgs a;
  a.X = X; a.ldx = ldx; a.wblob = blob; a.bias = bias; a.fv = fv; a.ldfv = ldfv; a.sigma = sigma; a.np = np; a.sch = sch;
  const size_t smem = kXBytes + kHBytes + kNst * kStageBytes + kFusedLayers * 144 * sizeof(float);
  static bool attr_done = false;
  static int num_sms = 148;
  if (!attr_done) {
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_decoder_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SHERF_CUDA_OK(cudaFuncSetAttribute(k_decoder_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0;
    SHERF_CUDA_OK(cudaGetDevice(&dev));
    SHERF_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_done = true;
  }
  const int ntiles = (np + 127) / 128;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  if (prec == 3) k_decoder_fused<3><<<grid, 320, smem, st>>>(a);
  else k_decoder_fused<1><<<grid, 320, smem, st>>>(a);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}


int run_pack_fused_plan(const SherfWeights& w, unsigned char* blob, float* bias, FusedPlan& plan, cudaStream_t st) {
  int rc = run_pack_fused(w, blob, bias, plan.sch, st);
  if (rc) return rc;
  plan.blob = blob;
  plan.bias = bias;
  return SHERF_OK;
}

int run_decoder_fused_plan(int prec, const FusedPlan& plan, const float* X, int ldx, float* fv, int ldfv, float* sigma, int np, cudaStream_t st) {
  return run_decoder_fused(prec, plan.sch, plan.blob, plan.bias, X, ldx, fv, ldfv, sigma, np, st);
}

}  // namespace sherf
