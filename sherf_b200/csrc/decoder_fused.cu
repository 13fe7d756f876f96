// Fused NeRF decoder on the tensor cores: pts_linears[0..7] (with the skip concat), feature_linear + alpha_linear,
// views_linear (feature | PE4(dir) | tok1) and the rgb head (triplane.py:293-314) for 128-point tiles, one persistent CTA per SM.  Activations never leave the SM:
//   hi parts   : shared memory, UMMA K-major no-swizzle canonical layout (padded LBO, see mlp_umma.cu)
//   lo parts   : tensor memory (3xTF32 error compensation), consumed by tcgen05.mma with the A operand in TMEM
//   accumulator: tensor memory, read back by the epilogue warps (bias + ReLU + tf32 split) straight into the next layer's operands
// Weights stream from L2 through a 3-stage shared-memory ring filled by a TMA (cp.async.bulk) producer thread.
// Warp roles: warps 0-7 = tile loader + epilogue (warp w owns TMEM lane quarter w & 3, column half w >> 2),
//             warp 8 lane 0 = MMA issuer, warp 9 lane 0 = TMA weight producer.
#include "common.cuh"
#include "stages.cuh"
#include "umma.cuh"
#include <cstdlib>

namespace sherf {

// cycle-counter tracing (sherf_debug_set_trace): compile with -DSHERF_FUSED_TRACE to enable
#ifdef SHERF_FUSED_TRACE
#define TRACE_CLK() clock64()
#else
#define TRACE_CLK() 0LL
#endif

constexpr int kFusedChunks = 44;
constexpr int kFusedLayers = 10;
constexpr int kNst = 3;                                  // weight ring stages (32 k-columns each; a 6 x 16 ring measured slower)
constexpr uint32_t kLbo = 2064;                          // padded K-direction stride of A operands (bytes)
constexpr uint32_t kXBytes = 18 * kLbo, kHBytes = 32 * kLbo;
constexpr uint32_t kStageBytes = 2 * 8 * 144 * 16;       // hi + lo, 8 core-matrix columns, up to 144 rows
constexpr uint32_t kColD0 = 0, kColD1 = 144, kColHlo = 288, kColXlo = 416;   // 2 x 144 accumulator + 128 H_lo + 72 X_lo = 488 <= 512

static_assert(kFusedChunks == 44 && kFusedLayers == 10, "FusedSchedule (stages.cuh) is sized for 44 chunks / 10 layers");

struct FusedArgs {
  const float* X; int ldx;             // [np][72] decoder input rows (PE6(can) | tok0 | 0)
  const unsigned char* wblob;          // packed chunks in schedule order
  const float* bias;                   // [9][144] (row 8: feature bias 0..127, alpha bias at 128)
  const float* fv; int ldfv;           // fv[:, 128:188] = PE4(cdir) | tok1 | 0 : the non-feature part of views_linear's input
  float* sigma;                        // [np]
  float* rgb;                          // [np][3]
  const float* rgb_w; const float* rgb_b;   // rgb_linear [3][64], [3]
  int np;
  long long* trace;                    // optional [gridDim][16] cycle counters (diagnostics)
  int dbg_flags;                       // timing experiments only (SHERF_FUSED_DBG): 1 skip proxy fence, 2 skip st wait
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
  __shared__ __align__(8) uint64_t full_bar[kNst], empty_bar[kNst], acc_bar, x_bar, hchunk_bar[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ FusedSchedule s_sch;
  __shared__ float s_rgbw[3 * 64 + 4];
  __shared__ float s_part[128 * 3];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (int)(sizeof(FusedSchedule) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&s_sch)[i] = reinterpret_cast<const uint32_t*>(&a.sch)[i];

  if (tid == 0) {
    for (int s = 0; s < kNst; ++s) { umma::mbar_init(&full_bar[s], 1); umma::mbar_init(&empty_bar[s], 1); }
    umma::mbar_init(&acc_bar, 1);
    umma::mbar_init(&x_bar, 256);
    for (int j = 0; j < 4; ++j) umma::mbar_init(&hchunk_bar[j], 256);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < kFusedLayers * 144; i += blockDim.x) s_bias[i] = a.bias[i];
  for (int i = tid; i < 3 * 64 + 3; i += blockDim.x) s_rgbw[i] = i < 192 ? a.rgb_w[i] : a.rgb_b[i - 192];
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int ntiles = (a.np + 127) / 128;

  if (warp == 9) {
    // ===================== TMA weight producer =====================
    if (lane == 0) {
      uint32_t cc = 0;
      long long t_wait = 0, t0 = TRACE_CLK();
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int c = 0; c < kFusedChunks; ++c, ++cc) {
          const int s = cc % kNst;
          const long long w0 = TRACE_CLK();
          umma::mbar_wait(&empty_bar[s], ((cc / kNst) & 1) ^ 1);
          t_wait += TRACE_CLK() - w0;
          const FusedChunk ch = s_sch.ch[c];
          const uint32_t bytes = (PREC == 3) ? ch.w_bytes : ch.w_bytes / 2;      // single-pass TF32 needs the hi half only
          umma::mbar_arrive_expect_tx(&full_bar[s], bytes);
          umma::bulk_g2s(Wst + s * kStageBytes, a.wblob + ch.w_off, bytes, &full_bar[s]);
        }
      }
      if (a.trace) { a.trace[blockIdx.x * 16 + 0] = t_wait; a.trace[blockIdx.x * 16 + 1] = TRACE_CLK() - t0; }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    {
      uint32_t cc = 0, par_x = 0, par_h[4] = {0, 0, 0, 0};
      const uint32_t el = umma::elect_one();     // one lane issues every MMA / commit of this warp
      long long t_op = 0, t_full = 0, t0 = TRACE_CLK();
      const uint32_t x_s = umma::smem_u32(X_hi), h_s = umma::smem_u32(H_hi), w_s = umma::smem_u32(Wst);
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int c = 0; c < kFusedChunks; ++c, ++cc) {
          const FusedChunk ch = s_sch.ch[c];             // (reading the schedule from the kernel-parameter bank measured 6 % slower)
          // operand readiness: X is loaded once per tile; H arrives from the previous layer's epilogue in 32-column chunks
          const long long w0 = TRACE_CLK();
          if (ch.src == 0) {
            if (ch.layer == 0 && ch.first) { umma::mbar_wait(&x_bar, par_x); par_x ^= 1; }
          } else if ((ch.kg0 & 7) == 0) {                   // first weight chunk that touches a new 32-column block of H
            const int j = ch.kg0 / 8;
            umma::mbar_wait(&hchunk_bar[j], par_h[j]);
            par_h[j] ^= 1;
          }
          const int s = cc % kNst;
          const long long w1 = TRACE_CLK();
          umma::mbar_wait(&full_bar[s], (cc / kNst) & 1);
          t_op += w1 - w0;
          t_full += TRACE_CLK() - w1;
          umma::tc_fence_after_sync();
          const uint32_t Np = s_sch.layer_np[ch.layer];
          const uint32_t idesc = umma::make_idesc_tf32(128, (int)Np);
          const uint32_t d_col = tmem_base + ((ch.layer & 1) ? kColD1 : kColD0);      // accumulators alternate per layer
          const uint32_t w_lbo = Np * 16u;
          const uint32_t w_hi = w_s + (uint32_t)s * kStageBytes, w_lo = w_hi + (uint32_t)ch.nkg * w_lbo;
          // descriptors of MMA k-step 0; later steps only advance the 14-bit start-address field (no carry: smem < 256 KB)
          const uint64_t ah0 = umma::make_smem_desc((ch.src == 0 ? x_s : h_s) + (uint32_t)ch.kg0 * kLbo, kLbo, 128u);
          const uint64_t wh0 = umma::make_smem_desc(w_hi, w_lbo, 128u);
          const uint64_t wl0 = umma::make_smem_desc(w_lo, w_lbo, 128u);
          const uint32_t alo0 = tmem_base + (ch.src == 0 ? kColXlo : kColHlo) + (uint32_t)ch.kg0 * 4u;
          const uint64_t da = (uint64_t)((2u * kLbo) >> 4), dw = (uint64_t)((2u * w_lbo) >> 4);
          const int nsteps = ch.nkg >> 1;
          const uint32_t acc0 = ch.first ? 0u : 1u;
#pragma unroll
          for (int st = 0; st < 4; ++st) {
            if (st < nsteps) {
              const uint32_t acc = st == 0 ? acc0 : 1u;
              if (PREC == 3) {
                umma::mma_tf32_ts_e(d_col, alo0 + (uint32_t)st * 8u, wh0 + (uint64_t)st * dw, idesc, acc, el);
                umma::mma_tf32_ss_e(d_col, ah0 + (uint64_t)st * da, wl0 + (uint64_t)st * dw, idesc, 1u, el);
                umma::mma_tf32_ss_e(d_col, ah0 + (uint64_t)st * da, wh0 + (uint64_t)st * dw, idesc, 1u, el);
              } else {
                umma::mma_tf32_ss_e(d_col, ah0 + (uint64_t)st * da, wh0 + (uint64_t)st * dw, idesc, acc, el);
              }
            }
          }
          umma::mma_commit_e(&empty_bar[s], el);                 // weight stage reusable once these MMAs retire
          if (ch.last) umma::mma_commit_e(&acc_bar, el);         // layer accumulator complete
          __syncwarp();
        }
      }
      if (a.trace && lane == 0) { a.trace[blockIdx.x * 16 + 2] = t_op; a.trace[blockIdx.x * 16 + 3] = t_full; a.trace[blockIdx.x * 16 + 4] = TRACE_CLK() - t0; }
    }
  } else {
    // ===================== tile loader + epilogue (warps 0-7) =====================
    const int q = warp & 3, hsel = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    uint32_t par_acc = 0;
    long long t_acc = 0, t_x = 0, t0 = TRACE_CLK();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long x0 = TRACE_CLK();
      const int m = tile * 128 + row;
      const bool row_ok = m < a.np;
      // ---- X tile: 72 columns = 18 core-matrix columns; this thread owns columns [36*hsel, 36*hsel+36) of its row ----
      {
        const float4* xr = reinterpret_cast<const float4*>(a.X + (size_t)m * a.ldx + 36 * hsel);
        float4 v[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = row_ok ? __ldg(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t lo[36];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const float4 h = make_float4(umma::to_tf32(v[i].x), umma::to_tf32(v[i].y), umma::to_tf32(v[i].z), umma::to_tf32(v[i].w));
          *reinterpret_cast<float4*>(X_hi + (9 * hsel + i) * kLbo + row * 16) = h;
          lo[4 * i + 0] = __float_as_uint(umma::to_tf32(v[i].x - h.x));
          lo[4 * i + 1] = __float_as_uint(umma::to_tf32(v[i].y - h.y));
          lo[4 * i + 2] = __float_as_uint(umma::to_tf32(v[i].z - h.z));
          lo[4 * i + 3] = __float_as_uint(umma::to_tf32(v[i].w - h.w));
        }
        if (PREC == 3) {
          const uint32_t xlo = tmem_base + lane_base + kColXlo + (uint32_t)(36 * hsel);
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const uint32_t t8[8] = {lo[8 * g8], lo[8 * g8 + 1], lo[8 * g8 + 2], lo[8 * g8 + 3], lo[8 * g8 + 4], lo[8 * g8 + 5], lo[8 * g8 + 6], lo[8 * g8 + 7]};
            umma::tmem_st8(xlo + (uint32_t)(8 * g8), t8);
          }
          const uint32_t t4[4] = {lo[32], lo[33], lo[34], lo[35]};
          umma::tmem_st4(xlo + 32u, t4);
          umma::tmem_st_wait();
        }
      }
      umma::fence_proxy_async_smem();
      umma::tc_fence_before_sync();
      mbar_arrive(&x_bar);
      t_x += TRACE_CLK() - x0;

      for (int l = 0; l < kFusedLayers; ++l) {
        const long long w0 = TRACE_CLK();
        umma::mbar_wait(&acc_bar, par_acc);
        t_acc += TRACE_CLK() - w0;
        par_acc ^= 1;
        umma::tc_fence_after_sync();
        const float* bl = s_bias + l * 144;
        const uint32_t d_col = tmem_base + lane_base + ((l & 1) ? kColD1 : kColD0);
        // all 8 warps sweep the accumulator in 32-column chunks (warp: lane quarter q, 16-column half hsel); after each chunk
        // the next layer's MMA may consume those 32 k-columns of H while this epilogue continues with the next chunk
        const int nchunk = (l == kFusedLayers - 1) ? 2 : 4;        // views_linear has 64 outputs, every other layer 128 (+ alpha)
        float pr = 0.f, pg = 0.f, pb = 0.f;                         // partial rgb_linear dots (last layer)
        uint32_t vnext[16];
        umma::tmem_ld16(d_col + (uint32_t)(16 * hsel), vnext);
        for (int j = 0; j < nchunk; ++j) {
          const int c0 = 32 * j + 16 * hsel;
          uint32_t v[16];
          umma::tmem_ld_wait();                             // chunk j has landed (and the previous chunk's tcgen05.st retired below)
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = vnext[i];
          if (j + 1 < nchunk) umma::tmem_ld16(d_col + (uint32_t)(c0 + 32), vnext);     // prefetch chunk j+1 while chunk j is processed
          if (l < kFusedLayers - 1) {
            // pts_linears: bias + ReLU; feature_linear (l == 8): bias only.  Result becomes the next layer's A operand.
            float x[16];
            uint32_t lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              x[i] = __uint_as_float(v[i]) + bl[c0 + i];
              if (l < 8) x[i] = fmaxf(x[i], 0.f);
              const float h = umma::to_tf32(x[i]);
              lo[i] = __float_as_uint(umma::to_tf32(x[i] - h));
              x[i] = h;
            }
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4)
              *reinterpret_cast<float4*>(H_hi + (c0 / 4 + g4) * kLbo + row * 16) = make_float4(x[4 * g4], x[4 * g4 + 1], x[4 * g4 + 2], x[4 * g4 + 3]);
            if (PREC == 3) {
              umma::tmem_st16(tmem_base + lane_base + kColHlo + (uint32_t)c0, lo);
              if (!(a.dbg_flags & 2)) umma::tmem_st_wait();
            }
            if (!(a.dbg_flags & 1)) umma::fence_proxy_async_smem();
            umma::tc_fence_before_sync();
            mbar_arrive(&hchunk_bar[j]);
          } else {
            // views_linear: bias + ReLU, then this thread's share of rgb_linear                          triplane.py:310-313
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float hv = fmaxf(__uint_as_float(v[i]) + bl[c0 + i], 0.f);
              pr = fmaf(hv, s_rgbw[c0 + i], pr); pg = fmaf(hv, s_rgbw[64 + c0 + i], pg); pb = fmaf(hv, s_rgbw[128 + c0 + i], pb);
            }
          }
        }
        if (l == 5) {
          // X is dead (its last reader, layer 5, has completed): load V = [PE4(cdir) | tok1 | 0] (60 -> 64 columns) into X's buffers
          const float4* vr = reinterpret_cast<const float4*>(a.fv + (size_t)m * a.ldfv + 128 + 32 * hsel);
          float vv[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = row_ok && (hsel == 0 || i < 7);          // columns 60..63 are padding
            const float4 f = ok ? __ldg(vr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            vv[4 * i] = f.x; vv[4 * i + 1] = f.y; vv[4 * i + 2] = f.z; vv[4 * i + 3] = f.w;
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t lo[16];
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              float h4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float xv = vv[half * 16 + g4 * 4 + e];
                h4[e] = umma::to_tf32(xv);
                lo[g4 * 4 + e] = __float_as_uint(umma::to_tf32(xv - h4[e]));
              }
              *reinterpret_cast<float4*>(X_hi + (8 * hsel + 4 * half + g4) * kLbo + row * 16) = make_float4(h4[0], h4[1], h4[2], h4[3]);
            }
            if (PREC == 3) umma::tmem_st16(tmem_base + lane_base + kColXlo + (uint32_t)(32 * hsel + 16 * half), lo);
          }
          if (PREC == 3) umma::tmem_st_wait();
          umma::fence_proxy_async_smem();
          umma::tc_fence_before_sync();
        }
        if (l == 8 && hsel == 0) {
          // alpha_linear = row 128 of the stacked feature/alpha weight                                  triplane.py:302
          uint32_t v8[8];
          umma::tmem_ld8(d_col + 128u, v8);
          umma::tmem_ld_wait();
          if (row_ok) a.sigma[m] = __uint_as_float(v8[0]) + bl[128];
        }
        if (l == kFusedLayers - 1) {
          // combine the two column halves of the row and apply the sigmoid clamp                       triplane.py:313-314
          if (hsel == 1) { s_part[row * 3] = pr; s_part[row * 3 + 1] = pg; s_part[row * 3 + 2] = pb; }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (hsel == 0 && row_ok) {
            const float zr = pr + s_part[row * 3] + s_rgbw[192], zg = pg + s_part[row * 3 + 1] + s_rgbw[193], zb = pb + s_part[row * 3 + 2] + s_rgbw[194];
            a.rgb[(size_t)m * 3] = (1.f / (1.f + expf(-zr))) * (1.f + 2.f * 0.001f) - 0.001f;
            a.rgb[(size_t)m * 3 + 1] = (1.f / (1.f + expf(-zg))) * (1.f + 2.f * 0.001f) - 0.001f;
            a.rgb[(size_t)m * 3 + 2] = (1.f / (1.f + expf(-zb))) * (1.f + 2.f * 0.001f) - 0.001f;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");           // s_part is reused by the next tile
        }
      }
    }
    if (a.trace && tid == 0) { a.trace[blockIdx.x * 16 + 5] = t_acc; a.trace[blockIdx.x * 16 + 6] = t_x; a.trace[blockIdx.x * 16 + 7] = TRACE_CLK() - t0; }
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
  const int l = blockIdx.x, n = threadIdx.x;     // 10 x 144
  float v = 0.f;
  if (l < 8) { if (n < 128) v = w.pts_b[l][n]; }
  else if (l == 8) { if (n < 128) v = w.feature_b[n]; else if (n == 128) v = w.alpha_b[0]; }
  else { if (n < 64) v = w.views_b[n]; }
  bias[l * 144 + n] = v;
}

long long* g_fused_trace = nullptr;   // set by sherf_debug_set_trace (diagnostics)

size_t fused_blob_bytes() { return (size_t)2 * (128 * (72 + 128 * 4 + 200 + 128 * 2) + 144 * 128 + 64 * 192) * 4 + 1024; }

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
  seg(9, w.views_w, nullptr, 64, 64, 187, 0, 128, 1, true, false);        // views_linear input = [feature | PE4(dir) | tok1]  triplane.py:308
  seg(9, w.views_w, nullptr, 64, 64, 187, 128, 59, 0, false, true);       // the 59 non-feature columns come from X's (reused) buffers
  if (c != kFusedChunks) { set_error("internal: fused schedule has %d chunks", c); return SHERF_E_INVALID; }
  for (int l = 0; l < 8; ++l) sch.layer_np[l] = 128;
  sch.layer_np[8] = 144;
  sch.layer_np[9] = 64;
  sch.pad[0] = sch.pad[1] = 0;
  if (!g_pack_plan_only) {
    k_pack_fused<<<dim3(8, kFusedChunks), 256, 0, st>>>(jobs, blob);
    SHERF_LAUNCH_CHECK();
    k_fused_bias<<<kFusedLayers, 144, 0, st>>>(w, bias);
    SHERF_LAUNCH_CHECK();
  }
  return SHERF_OK;
}

int run_decoder_fused(int prec, const FusedSchedule& sch, const unsigned char* blob, const float* bias, const float* X, int ldx, const float* fv,
                      int ldfv, float* sigma, float* rgb, const float* rgb_w, const float* rgb_b, int np, cudaStream_t st) {
  if (np <= 0) return SHERF_OK;
  FusedArgs a;
  a.X = X; a.ldx = ldx; a.wblob = blob; a.bias = bias; a.fv = fv; a.ldfv = ldfv; a.sigma = sigma; a.rgb = rgb; a.rgb_w = rgb_w; a.rgb_b = rgb_b; a.np = np; a.sch = sch;
  a.trace = g_fused_trace;
  { const char* e = getenv("SHERF_FUSED_DBG"); a.dbg_flags = e ? atoi(e) : 0; }
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

int run_decoder_fused_plan(int prec, const FusedPlan& plan, const float* X, int ldx, const float* fv, int ldfv, float* sigma, float* rgb,
                           const float* rgb_w, const float* rgb_b, int np, cudaStream_t st) {
  return run_decoder_fused(prec, plan.sch, plan.blob, plan.bias, X, ldx, fv, ldfv, sigma, rgb, rgb_w, rgb_b, np, st);
}

}  // namespace sherf
