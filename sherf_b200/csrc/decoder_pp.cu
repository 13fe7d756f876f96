// NeRF decoder (triplane.py:285-316) on the tensor cores, "ping-pong" edition: bf16x3 split products, two 128-point
// tiles in flight per SM, every activation in tensor memory.
//
//   arithmetic : a = a_hi + a_lo, w = w_hi + w_lo with bf16 parts (16 significand bits kept); a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
//                in three kind::f16 MMAs with fp32 accumulation.  Half the tensor-core time of the 3xTF32 kernel
//                (decoder_fused.cu) at a per-product error of ~2^-16, which leaves the rendered image at the same distance
//                from the reference as fp32 arithmetic does (tests/test_parity_gpu.py).
//   slots      : TMEM is split in two 256-column slots {D 128 | H_hi 64 | H_lo 64}; slot A's layer-l epilogue (TMEM ->
//                bias/ReLU/split -> TMEM, 4 warps, one thread per point) runs while the tensor pipe works on slot B's layer l.
//   weights    : streamed once per tile PAIR through a 4 x 32 KB TMA ring; every stage is consumed by both slots before it
//                is released (L2 -> SM weight traffic is the second bound of this kernel).
//   inputs     : X = [PE6(can) | tok0] and V = [PE4(dir) | tok1] arrive as packed bf16 hi/lo tiles in UMMA canonical layout
//                (written by k_pack_xv / the transformer kernel) and are bulk-copied into shared memory; they are the only
//                A operands read from shared memory.
// Warp roles: 0-3 epilogue of slot 0, 4-7 epilogue of slot 1, 8 MMA issuer, 9 weight producer, 10 X/V producer.
#include "common.cuh"
#include "stages.cuh"
#include "umma.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

namespace sherf {
namespace dpp {

// cycle-counter tracing (sherf_debug_set_trace): compile with -DSHERF_FUSED_TRACE to enable
#ifdef SHERF_FUSED_TRACE
#define TRACE_CLK() clock64()
#else
#define TRACE_CLK() 0LL
#endif

constexpr int kLayers = 10;
constexpr int kChunks = 23;
constexpr int kStages = 4;
constexpr uint32_t kStageBytes = 32768;                 // 4 k-steps x (hi + lo) x 128 rows x 32 B
constexpr uint32_t kXTile = 40960, kVTile = 32768;      // [hi | lo] x [kg][128 rows][8 bf16]
constexpr uint32_t kColD = 0, kColHhi = 128, kColHlo = 192, kSlotCols = 256;

// k-steps (16 k-columns) per layer and the A operand each one reads:
//   layer 0: X[0..4]; layers 1-4, 6-8: H[0..7]; layer 5: X[0..4] then H[0..7] (skip concat, triplane.py:299-300);
//   layer 9 (views_linear): H[0..7] (= feature) then V[0..3]
__host__ __device__ constexpr int layer_nks(int l) { return l == 0 ? 5 : l == 5 ? 13 : l == 9 ? 12 : 8; }
__host__ __device__ constexpr int layer_n(int l) { return l == 9 ? 64 : 128; }
__host__ __device__ constexpr int layer_nchunks(int l) { return (layer_nks(l) + 3) / 4; }

struct Args {
  const unsigned char* xp;     // [ntiles][kXTile]
  const unsigned char* vp;     // [ntiles][kVTile]
  const unsigned char* wblob;  // chunks in schedule order
  uint32_t w_off[kChunks];
  const float* bias;           // [10][128]
  const float* alpha_w; const float* alpha_b;   // alpha_linear [128], [1]
  const float* rgb_w; const float* rgb_b;       // rgb_linear [3][64], [3]
  float* sigma;                // [np]
  float* rgb;                  // [np][3]
  int np;
  DevCount dc;
  long long* trace;            // optional [gridDim][16] cycle counters (diagnostics)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}


// All MMAs of layer L for one slot: straight-line code, every operand offset a compile-time constant (a run-time layer /
// k-step dispatch measured ~340 issue cycles per k-step, five times the tensor time of its three MMAs).  L = 1 stands for
// every plain 128 -> 128 layer (1-4, 6-8).
template <int L>
__device__ __forceinline__ void issue_layer(const uint32_t slot, const uint32_t xbuf, const uint32_t w_s, const uint32_t gc, const bool release,
                                            uint64_t* full_bar, uint64_t* empty_bar, uint64_t* xfull, uint32_t& par_x, const uint32_t el) {
  constexpr int NKS = layer_nks(L), N = layer_n(L), NCH = (NKS + 3) / 4;
  constexpr uint32_t idesc = umma::make_idesc_bf16(128, N);
  constexpr uint32_t w_lbo = (uint32_t)N * 16u;
  const uint64_t xd = umma::make_smem_desc(xbuf, 2048u, 128u);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const uint32_t g = gc + (uint32_t)c;
    const uint32_t st = g % kStages;
    umma::mbar_wait(&full_bar[st], (g / kStages) & 1);
    if (L == 9 && c == 2) { umma::mbar_wait(xfull, par_x); par_x ^= 1; }       // V tile landed
    umma::tc_fence_after_sync();
    const int nk = NKS - 4 * c < 4 ? NKS - 4 * c : 4;
    const uint64_t wd = umma::make_smem_desc(w_s + st * kStageBytes, w_lbo, 128u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < nk) {
        const int ks = 4 * c + j;
        const uint64_t bh = wd + (uint64_t)(((uint32_t)j * 2u * w_lbo) >> 4), bl = wd + (uint64_t)(((uint32_t)(nk + j) * 2u * w_lbo) >> 4);
        const uint32_t acc = ks == 0 ? 0u : 1u;
        const bool from_h = L == 0 ? false : L == 5 ? ks >= 5 : L == 9 ? ks < 8 : true;
        const int sk = L == 5 ? (ks < 5 ? ks : ks - 5) : L == 9 ? (ks < 8 ? ks : ks - 8) : ks;
        if (from_h) {
          const uint32_t ah = slot + kColHhi + (uint32_t)sk * 8u, al = slot + kColHlo + (uint32_t)sk * 8u;
          umma::mma_bf16_ts_e(slot + kColD, ah, bh, idesc, acc, el);
          umma::mma_bf16_ts_e(slot + kColD, al, bh, idesc, 1u, el);
          umma::mma_bf16_ts_e(slot + kColD, ah, bl, idesc, 1u, el);
        } else {
          constexpr uint32_t lo_off = (L == 9) ? kVTile / 2 : kXTile / 2;
          const uint64_t ah = xd + (uint64_t)(((uint32_t)sk * 4096u) >> 4), al = xd + (uint64_t)((lo_off + (uint32_t)sk * 4096u) >> 4);
          umma::mma_bf16_ss_e(slot + kColD, ah, bh, idesc, acc, el);
          umma::mma_bf16_ss_e(slot + kColD, al, bh, idesc, 1u, el);
          umma::mma_bf16_ss_e(slot + kColD, ah, bl, idesc, 1u, el);
        }
      }
    }
    if (release) umma::mma_commit_e(&empty_bar[st], el);
  }
}

// One layer's epilogue for one point (thread = accumulator row): TMEM accumulator -> bias (+ReLU) -> bf16 hi/lo split -> TMEM
// operand of the next layer.  MODE 0: pts_linears (bias + ReLU); 1: pts_linears[7], additionally alpha_linear's dot product
// (triplane.py:302); 2: feature_linear (bias only); 3: views_linear (bias + ReLU) followed by rgb_linear's three dot products.
template <int MODE>
__device__ __forceinline__ void epi_layer(const uint32_t slot, const float* __restrict__ bl, const float* __restrict__ s_alpha,
                                          const float* __restrict__ s_rgbw, float& o0, float& o1, float& o2) {
  constexpr int NBLK = MODE == 3 ? 2 : 4;                     // 32-column blocks of the accumulator
  float e[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint32_t v[NBLK][32];
  umma::tmem_ld32(slot + kColD, v[0]);
#pragma unroll
  for (int j = 0; j < NBLK; ++j) {
    const int c0 = 32 * j;
    float b[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 f = *reinterpret_cast<const float4*>(bl + c0 + 4 * i);
      b[4 * i] = f.x; b[4 * i + 1] = f.y; b[4 * i + 2] = f.z; b[4 * i + 3] = f.w;
    }
    umma::tmem_ld_wait();
    if (j + 1 < NBLK) umma::tmem_ld32(slot + kColD + (uint32_t)(32 * (j + 1)), v[j + 1]);
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      x[i] = __uint_as_float(v[j][i]) + b[i];
      if (MODE != 2) x[i] = fmaxf(x[i], 0.f);
    }
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(s_alpha + c0 + 4 * i);
        e[0] = fmaf(x[4 * i], w.x, e[0]); e[1] = fmaf(x[4 * i + 1], w.y, e[1]); e[2] = fmaf(x[4 * i + 2], w.z, e[2]); e[3] = fmaf(x[4 * i + 3], w.w, e[3]);
      }
    }
    if (MODE == 3) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w = *reinterpret_cast<const float4*>(s_rgbw + 64 * ch + c0 + 4 * i);
          e[2 * ch] = fmaf(x[4 * i], w.x, e[2 * ch]); e[2 * ch + 1] = fmaf(x[4 * i + 1], w.y, e[2 * ch + 1]);
          e[2 * ch] = fmaf(x[4 * i + 2], w.z, e[2 * ch]); e[2 * ch + 1] = fmaf(x[4 * i + 3], w.w, e[2 * ch + 1]);
        }
      }
    } else {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) umma::split_bf16x2(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
      umma::tmem_st16(slot + kColHhi + (uint32_t)(c0 / 2), hi);
      umma::tmem_st16(slot + kColHlo + (uint32_t)(c0 / 2), lo);
    }
  }
  if (MODE != 3) umma::tmem_st_wait();
  if (MODE == 1) o0 = (e[0] + e[1]) + (e[2] + e[3]);
  if (MODE == 3) { o0 = e[0] + e[1]; o1 = e[2] + e[3]; o2 = e[4] + e[5]; }
}

__global__ void __launch_bounds__(352, 1) k_decoder_pp(const Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* XV = smem;                                   // 2 slots x kXTile
  unsigned char* Wst = smem + 2 * kXTile;                     // kStages x kStageBytes
  float* s_bias = reinterpret_cast<float*>(Wst + kStages * kStageBytes);   // [10][128]
  float* s_alpha = s_bias + kLayers * 128;                    // [128] + bias
  float* s_rgbw = s_alpha + 132;                              // [3][64] + 3
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_bar[2], h_bar[2], xfull_bar[2], xfree_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { umma::mbar_init(&full_bar[s], 1); umma::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(&acc_bar[s], 1); umma::mbar_init(&h_bar[s], 128);
      umma::mbar_init(&xfull_bar[s], 1); umma::mbar_init(&xfree_bar[s], 1);
    }
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < kLayers * 128; i += blockDim.x) s_bias[i] = a.bias[i];
  for (int i = tid; i < 129; i += blockDim.x) s_alpha[i] = i < 128 ? a.alpha_w[i] : a.alpha_b[0];
  for (int i = tid; i < 195; i += blockDim.x) s_rgbw[i] = i < 192 ? a.rgb_w[i] : a.rgb_b[i - 192];
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int np = resolve_np(a.np, a.dc);
  const int ntiles = (np + 127) / 128;
  const int G = gridDim.x;
  // tile pair p of this CTA: slot 0 <- tile (2p)G + b, slot 1 <- tile (2p+1)G + b

  if (warp == 9) {
    // ===================== weight producer: one pass over the 23 chunks per tile pair =====================
    if (lane == 0) {
      uint32_t gc = 0;
      long long t_wait = 0, t0c = TRACE_CLK();
      for (int t0 = blockIdx.x; t0 < ntiles; t0 += 2 * G) {
        int c = 0;
        for (int l = 0; l < kLayers; ++l) {
          const int nks = layer_nks(l), N = layer_n(l);
          for (int k0 = 0; k0 < nks; k0 += 4, ++c, ++gc) {
            const int s = gc % kStages;
            const uint32_t bytes = (uint32_t)((nks - k0 < 4 ? nks - k0 : 4) * N * 64);
            const long long w0 = TRACE_CLK();
            umma::mbar_wait(&empty_bar[s], ((gc / kStages) & 1) ^ 1);
            t_wait += TRACE_CLK() - w0;
            umma::mbar_arrive_expect_tx(&full_bar[s], bytes);
            umma::bulk_g2s(Wst + s * kStageBytes, a.wblob + a.w_off[c], bytes, &full_bar[s]);
          }
        }
      }
      if (a.trace) { a.trace[blockIdx.x * 16 + 0] = t_wait; a.trace[blockIdx.x * 16 + 1] = TRACE_CLK() - t0c; }
    }
  } else if (warp == 10) {
    // ===================== X / V producer =====================
    // Slot s's buffer holds X from before layer 0 until layer 5 has retired, then V until layer 9 has retired.  The events
    // it waits for happen in exactly the order it waits (A5 < B5 < A9 < B9), so one in-order thread cannot deadlock.
    if (lane == 0) {
      uint32_t nfree[2] = {0, 0};
      for (int t0 = blockIdx.x; t0 < ntiles; t0 += 2 * G) {
        const int nact = (t0 + G < ntiles) ? 2 : 1;
        for (int s = 0; s < nact; ++s) {
          const int tile = t0 + s * G;
          if (t0 != (int)blockIdx.x) { umma::mbar_wait(&xfree_bar[s], nfree[s] & 1); ++nfree[s]; }     // layer 9 of the previous tile retired
          umma::mbar_arrive_expect_tx(&xfull_bar[s], kXTile);
          umma::bulk_g2s(XV + s * kXTile, a.xp + (size_t)tile * kXTile, kXTile, &xfull_bar[s]);
        }
        for (int s = 0; s < nact; ++s) {
          const int tile = t0 + s * G;
          umma::mbar_wait(&xfree_bar[s], nfree[s] & 1); ++nfree[s];                                     // layer 5 retired
          umma::mbar_arrive_expect_tx(&xfull_bar[s], kVTile);
          umma::bulk_g2s(XV + s * kXTile, a.vp + (size_t)tile * kVTile, kVTile, &xfull_bar[s]);
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    uint32_t gc = 0, par_h[2] = {0, 0}, par_x[2] = {0, 0};
    long long t_h = 0, t_x = 0, t_mma = 0, t_commit = 0, t0c = TRACE_CLK();
    const uint32_t xv_s = umma::smem_u32(XV), w_s = umma::smem_u32(Wst);
    const uint32_t el = umma::elect_one();                        // the warp stays converged; one lane issues everything
    for (int t0 = blockIdx.x; t0 < ntiles; t0 += 2 * G) {
      const int nact = (t0 + G < ntiles) ? 2 : 1;
      const bool first_pair = t0 == (int)blockIdx.x;
      for (int l = 0; l < kLayers; ++l) {
        const int nch = layer_nchunks(l);
        for (int s = 0; s < nact; ++s) {
          const uint32_t slot = tmem_base + (uint32_t)s * kSlotCols;
          const uint32_t xbuf = xv_s + (uint32_t)s * kXTile;
          // accumulator free (previous epilogue of this slot has read it) and, for l > 0, H complete
          const long long w0 = TRACE_CLK();
          if (!(first_pair && l == 0)) { umma::mbar_wait(&h_bar[s], par_h[s]); par_h[s] ^= 1; }
          t_h += TRACE_CLK() - w0;
          const long long w2 = TRACE_CLK();
          if (l == 0) { umma::mbar_wait(&xfull_bar[s], par_x[s]); par_x[s] ^= 1; }
          t_x += TRACE_CLK() - w2;
          const long long w3 = TRACE_CLK();
          const bool release = s == nact - 1;                      // both slots have consumed the layer's weight stages
          switch (l) {
            case 0: issue_layer<0>(slot, xbuf, w_s, gc, release, full_bar, empty_bar, &xfull_bar[s], par_x[s], el); break;
            case 5: issue_layer<5>(slot, xbuf, w_s, gc, release, full_bar, empty_bar, &xfull_bar[s], par_x[s], el); break;
            case 9: issue_layer<9>(slot, xbuf, w_s, gc, release, full_bar, empty_bar, &xfull_bar[s], par_x[s], el); break;
            default: issue_layer<1>(slot, xbuf, w_s, gc, release, full_bar, empty_bar, &xfull_bar[s], par_x[s], el); break;
          }
          t_mma += TRACE_CLK() - w3;
          const long long w5 = TRACE_CLK();
          umma::mma_commit_e(&acc_bar[s], el);
          if (l == 5 || l == 9) umma::mma_commit_e(&xfree_bar[s], el);
          t_commit += TRACE_CLK() - w5;
        }
        gc += (uint32_t)nch;
      }
    }
    if (a.trace && lane == 0) { a.trace[blockIdx.x * 16 + 2] = t_h; a.trace[blockIdx.x * 16 + 3] = 0; a.trace[blockIdx.x * 16 + 4] = TRACE_CLK() - t0c;
      a.trace[blockIdx.x * 16 + 8] = t_x; a.trace[blockIdx.x * 16 + 9] = t_mma; a.trace[blockIdx.x * 16 + 10] = t_commit; }
  } else {
    // ===================== epilogue: warps 0-3 serve slot 0, warps 4-7 slot 1; one thread per point =====================
    const int s = warp >> 2, q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t slot = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)s * kSlotCols;
    uint32_t par_acc = 0;
    long long t_acc = 0, t0c = TRACE_CLK();
    for (int tile = blockIdx.x + s * G; tile < ntiles; tile += 2 * G) {
      const int m = tile * 128 + row;
      const bool row_ok = m < np;
      for (int l = 0; l < kLayers; ++l) {
        const long long w0 = TRACE_CLK();
        umma::mbar_wait(&acc_bar[s], par_acc);
        t_acc += TRACE_CLK() - w0;
        par_acc ^= 1;
        umma::tc_fence_after_sync();
        const float* bl = s_bias + l * 128;
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        // the four layer flavours are separate straight-line instantiations (a single body with run-time `l` tests measured
        // 2x the instructions and no cross-element scheduling)
        if (l == 9) epi_layer<3>(slot, bl, s_alpha, s_rgbw, o0, o1, o2);
        else if (l == 8) epi_layer<2>(slot, bl, s_alpha, s_rgbw, o0, o1, o2);
        else if (l == 7) epi_layer<1>(slot, bl, s_alpha, s_rgbw, o0, o1, o2);
        else epi_layer<0>(slot, bl, s_alpha, s_rgbw, o0, o1, o2);
        umma::tc_fence_before_sync();
        mbar_arrive(&h_bar[s]);                                     // H written (l < 9) and accumulator drained
        if (l == 7 && row_ok) a.sigma[m] = o0 + s_alpha[128];
        if (l == 9 && row_ok) {
          const float zr = o0 + s_rgbw[192], zg = o1 + s_rgbw[193], zb = o2 + s_rgbw[194];                 // triplane.py:313-314
          a.rgb[(size_t)m * 3] = (1.f / (1.f + expf(-zr))) * (1.f + 2.f * 0.001f) - 0.001f;
          a.rgb[(size_t)m * 3 + 1] = (1.f / (1.f + expf(-zg))) * (1.f + 2.f * 0.001f) - 0.001f;
          a.rgb[(size_t)m * 3 + 2] = (1.f / (1.f + expf(-zb))) * (1.f + 2.f * 0.001f) - 0.001f;
        }
      }
    }
    if (a.trace && tid == 0) { a.trace[blockIdx.x * 16 + 5] = t_acc; a.trace[blockIdx.x * 16 + 7] = TRACE_CLK() - t0c; }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight blob: chunk = [hi: nk x 2 core-matrix columns x N rows x 8 bf16][lo: same]; element (kg, n, e) of a chunk is
// W[n][wcol0[kg/2] + (kg&1)*8 + e] (zero beyond nvalid[kg/2] columns).
struct PackJob { const float* W; int ldw, N, nk; int wcol0[4], nvalid[4]; uint32_t off; };
struct PackJobs { PackJob j[kChunks]; };

__global__ void k_pack_pp(const PackJobs jobs, unsigned char* blob) {
  const PackJob jb = jobs.j[blockIdx.y];
  const int total = jb.nk * 2 * jb.N * 8;
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(blob + jb.off);
  __nv_bfloat16* lo = hi + total;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7, n = (i >> 3) % jb.N, kg = (i >> 3) / jb.N;
    const int j = kg >> 1, kk = (kg & 1) * 8 + e;
    const float v = kk < jb.nvalid[j] ? jb.W[(size_t)n * jb.ldw + jb.wcol0[j] + kk] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

__global__ void k_pp_bias(const SherfWeights w, float* bias) {
  const int l = blockIdx.x, n = threadIdx.x;     // 10 x 128
  float v = 0.f;
  if (l < 8) v = w.pts_b[l][n];
  else if (l == 8) v = w.feature_b[n];
  else if (n < 64) v = w.views_b[n];
  bias[l * 128 + n] = v;
}

// fp32 decoder inputs -> packed bf16 hi/lo tiles.  One block per 128-point tile, one thread per point.
__global__ void k_pack_xv(const float* __restrict__ x, int ldx, const float* __restrict__ fv, int ldfv, int np, unsigned char* __restrict__ xp,
                          unsigned char* __restrict__ vp) {
  const int tile = blockIdx.x, r = threadIdx.x, m = tile * 128 + r;
  const bool ok = m < np;
  uint4* xh = reinterpret_cast<uint4*>(xp + (size_t)tile * kXTile);
  uint4* vh = reinterpret_cast<uint4*>(vp + (size_t)tile * kVTile);
  for (int kg = 0; kg < 10; ++kg) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { const int c = kg * 8 + e; v[e] = (ok && c < 71) ? x[(size_t)m * ldx + c] : 0.f; }
    uint4 h, l;
    umma::split_bf16x2(v[0], v[1], h.x, l.x); umma::split_bf16x2(v[2], v[3], h.y, l.y);
    umma::split_bf16x2(v[4], v[5], h.z, l.z); umma::split_bf16x2(v[6], v[7], h.w, l.w);
    xh[kg * 128 + r] = h;
    xh[1280 + kg * 128 + r] = l;
  }
  for (int kg = 0; kg < 8; ++kg) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { const int c = kg * 8 + e; v[e] = (ok && c < 59) ? fv[(size_t)m * ldfv + 128 + c] : 0.f; }
    uint4 h, l;
    umma::split_bf16x2(v[0], v[1], h.x, l.x); umma::split_bf16x2(v[2], v[3], h.y, l.y);
    umma::split_bf16x2(v[4], v[5], h.z, l.z); umma::split_bf16x2(v[6], v[7], h.w, l.w);
    vh[kg * 128 + r] = h;
    vh[1024 + kg * 128 + r] = l;
  }
}

}  // namespace dpp

size_t pp_blob_bytes() {
  size_t b = 0;
  for (int l = 0; l < dpp::kLayers; ++l) b += (size_t)dpp::layer_nks(l) * dpp::layer_n(l) * 64;
  return b + 1024;
}
size_t pp_xv_bytes(int cap) { return (size_t)((cap + 127) / 128) * (dpp::kXTile + dpp::kVTile); }

int run_pack_pp(const SherfWeights& w, unsigned char* blob, float* bias, PpPlan& plan, cudaStream_t st) {
  dpp::PackJobs jobs;
  int c = 0;
  uint32_t off = 0;
  for (int l = 0; l < dpp::kLayers; ++l) {
    const float* W = l < 8 ? w.pts_w[l] : l == 8 ? w.feature_w : w.views_w;
    const int ldw = l == 0 ? 71 : l == 5 ? 199 : l == 9 ? 187 : 128;
    const int nks = dpp::layer_nks(l), N = dpp::layer_n(l);
    for (int k0 = 0; k0 < nks; k0 += 4, ++c) {
      dpp::PackJob& j = jobs.j[c];
      j.W = W; j.ldw = ldw; j.N = N; j.nk = nks - k0 < 4 ? nks - k0 : 4; j.off = off;
      for (int i = 0; i < 4; ++i) {
        const int ks = k0 + i;
        int col = 16 * ks, nv = 16;
        if (l == 0) { nv = ks == 4 ? 7 : 16; }
        else if (l == 5) { if (ks < 5) nv = ks == 4 ? 7 : 16; else col = 71 + 16 * (ks - 5); }     // cat([x, h])  triplane.py:299
        else if (l == 9) { if (ks == 11) nv = 11; }                                                 // [feature | PE4(dir) | tok1] = 187
        j.wcol0[i] = col; j.nvalid[i] = i < j.nk ? nv : 0;
      }
      plan.w_off[c] = off;
      off += (uint32_t)(j.nk * N * 64);
    }
  }
  if (c != dpp::kChunks) { set_error("internal: ping-pong decoder schedule has %d chunks", c); return SHERF_E_INVALID; }
  if (!g_pack_plan_only) {
    dpp::k_pack_pp<<<dim3(8, dpp::kChunks), 256, 0, st>>>(jobs, blob);
    SHERF_LAUNCH_CHECK();
    dpp::k_pp_bias<<<dpp::kLayers, 128, 0, st>>>(w, bias);
    SHERF_LAUNCH_CHECK();
  }
  plan.blob = blob;
  plan.bias = bias;
  return SHERF_OK;
}

int run_pack_xv(const float* x, int ldx, const float* fv, int ldfv, int np, unsigned char* xp, unsigned char* vp, cudaStream_t st) {
  if (np <= 0) return SHERF_OK;
  dpp::k_pack_xv<<<(np + 127) / 128, 128, 0, st>>>(x, ldx, fv, ldfv, np, xp, vp);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

int run_decoder_pp(const PpPlan& plan, const SherfWeights& w, const unsigned char* xp, const unsigned char* vp, float* sigma, float* rgb, int np,
                   cudaStream_t st, DevCount dc) {
  if (np <= 0) return SHERF_OK;
  dpp::Args a;
  a.xp = xp; a.vp = vp; a.wblob = plan.blob; a.bias = plan.bias;
  for (int c = 0; c < dpp::kChunks; ++c) a.w_off[c] = plan.w_off[c];
  a.alpha_w = w.alpha_w; a.alpha_b = w.alpha_b; a.rgb_w = w.rgb_w; a.rgb_b = w.rgb_b;
  a.sigma = sigma; a.rgb = rgb; a.np = np; a.dc = dc;
  a.trace = g_fused_trace;
  const size_t smem = 2 * dpp::kXTile + dpp::kStages * dpp::kStageBytes + (dpp::kLayers * 128 + 132 + 196) * sizeof(float);
  static bool attr_done = false;
  static int num_sms = 148;
  if (!attr_done) {
    SHERF_CUDA_OK(cudaFuncSetAttribute(dpp::k_decoder_pp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0;
    SHERF_CUDA_OK(cudaGetDevice(&dev));
    SHERF_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_done = true;
  }
  const int ntiles = (np + 127) / 128;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  dpp::k_decoder_pp<<<grid, 352, smem, st>>>(a);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

}  // namespace sherf
