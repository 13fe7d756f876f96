// Importance (fine) pass, SURVEY.md a13: renderer.py:373-393 in its repaired form (see include/sherf_b200.h).
//   k_importance_sample  : coarse ray-marcher weights (ray_marcher.py:25-50) -> sample_importance (renderer.py:483-501)
//                          -> sample_pdf (renderer.py:503-542) -> fine depths, one warp per ray
//   k_composite_merged   : unify_samples (renderer.py:446-456: concatenate + sort by depth) fused with the final ray march
//                          (ray_marcher.py:25-64) over the coarse and the fine compacted point lists, one warp per ray
// Both kernels are latency / shared-memory bound per-ray bookkeeping (28 B in + 20 B out per ray); the heavy
// part of the fine pass reuses the cull / gather / MLP stages on the fine point list.
#include "common.cuh"
#include "stages.cuh"

namespace sherf {

constexpr int kMaxS = 256;       // validate(): n_samples, n_importance <= 256

// ---------------------------------------------------------------------------------------------------------------------
// FROM_POINTS = true : weights come from the compacted coarse points (sigma of the MLP stage), culled samples have weight 0
// FROM_POINTS = false: weights are read from `w_in` [N*S] (sherf_debug_sample_importance)
template <bool FROM_POINTS>
__global__ void __launch_bounds__(128) k_importance_sample(const float* __restrict__ dirs, const float* __restrict__ nearv,
                                                           const float* __restrict__ farv, int N, int S, int SF,
                                                           const int* __restrict__ ray_start, const int* __restrict__ point_sample,
                                                           const float* __restrict__ sigma, const float* __restrict__ noise,
                                                           const float* __restrict__ w_in, const float* __restrict__ u,
                                                           float* __restrict__ t_fine, int* __restrict__ bins_out,
                                                           float* __restrict__ w_out) {
  __shared__ float s_w[4][kMaxS + 2];      // s_w[1 + i] = weight of sample i, s_w[0] = s_w[S + 1] = -inf (max_pool1d padding)
  __shared__ float s_c[4][kMaxS];          // smoothed weights, then the cdf (S - 1 entries)
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int n = blockIdx.x * 4 + wl;
  if (n >= N) return;
  float* w = s_w[wl];
  float* c = s_c[wl];
  const float nr = nearv[n], fr = farv[n];
  const float ninf = __int_as_float(0xff800000);
  for (int i = lane; i < S + 2; i += 32) w[i] = (i == 0 || i == S + 1) ? ninf : 0.f;
  __syncwarp();
  if (FROM_POINTS) {
    // ray_marcher.py:27-50 over the compacted survivors (same arithmetic as k_composite)
    const int b = ray_start[n], e = ray_start[n + 1];
    const float dx = dirs[n * 3], dy = dirs[n * 3 + 1], dz = dirs[n * 3 + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    float T = 1.f;
    for (int base = b; base < e; base += 32) {
      const int p = base + lane;
      float alpha = 0.f;
      int i = 0;
      if (p < e) {
        const int s = point_sample[p];
        i = s - n * S;
        const float t = sample_depth(nr, fr, i, S);
        const float delta = ((i == S - 1) ? 1e10f : (sample_depth(nr, fr, i + 1, S) - t)) * dnorm;
        float sg = sigma[p];
        if (noise) sg += noise[p];                                          // per SURVIVING point, compacted order (renderer.py:435-436)
        alpha = 1.f - expf(-(fmaxf(sg, 0.f) * delta));
      }
      const float f = (p < e) ? (1.f - alpha + 1e-10f) : 1.f;
      float incl = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= up;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.f;
      if (p < e) w[1 + i] = alpha * (T * excl);
      T *= __shfl_sync(0xffffffffu, incl, 31);
    }
  } else {
    for (int i = lane; i < S; i += 32) w[1 + i] = w_in[(size_t)n * S + i];
  }
  __syncwarp();
  if (w_out) for (int i = lane; i < S; i += 32) w_out[(size_t)n * S + i] = w[1 + i];
  // renderer.py:494-496: max_pool1d(k=2, s=1, pad=1) -> avg_pool1d(k=2, s=1) -> + 0.01
  //   m[j] = max(w[j-1], w[j]), j = 0..S ; a[j] = (m[j] + m[j+1]) / 2, j = 0..S-1
  float part = 0.f;
  for (int j = lane; j < S; j += 32) {
    const float m0 = fmaxf(w[j], w[j + 1]), m1 = fmaxf(w[j + 1], w[j + 2]);          // s_w is shifted by one
    const float a = __fadd_rn(__fmul_rn(__fadd_rn(m0, m1), 0.5f), 0.01f);
    // renderer.py:499 keeps a[1:-1]; sample_pdf :517 adds eps = 1e-5
    const float wk = __fadd_rn(a, 1e-5f);
    c[j] = wk;
    if (j >= 1 && j <= S - 2) part += wk;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  __syncwarp();
  // pdf = w / sum ; cdf = [0, cumsum(pdf)] (S - 1 entries) -- sequential like torch's CPU cumsum        renderer.py:518-521
  const int nb = S - 2;                                                             // N_samples_
  float pdf_reg[kMaxS / 32];
#pragma unroll
  for (int r = 0; r < kMaxS / 32; ++r) {
    const int k = r * 32 + lane;
    pdf_reg[r] = (k < nb) ? __fdiv_rn(c[k + 1], part) : 0.f;
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < kMaxS / 32; ++r) {
    const int k = r * 32 + lane;
    if (k < nb) w[k] = pdf_reg[r];                                                   // reuse s_w for the pdf
  }
  __syncwarp();
  if (lane == 0) {
    float acc = 0.f;
    c[0] = 0.f;
    for (int k = 0; k < nb; ++k) { acc = __fadd_rn(acc, w[k]); c[k + 1] = acc; }
  }
  __syncwarp();
  // inverse-CDF sampling                                                                                renderer.py:529-541
  const int ncdf = nb + 1;
  for (int j = lane; j < SF; j += 32) {
    const float uu = u[(size_t)n * SF + j];
    int lo = 0, hi = ncdf;                                                           // inds = #{cdf <= u}  (searchsorted right=True)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (c[mid] <= uu) lo = mid + 1; else hi = mid;
    }
    const int inds = lo;
    const int below = max(inds - 1, 0), above = min(inds, nb);
    const float c0 = c[below], c1 = c[above];
    const float b0 = __fmul_rn(0.5f, __fadd_rn(sample_depth(nr, fr, below, S), sample_depth(nr, fr, below + 1, S)));     // z_vals_mid :498
    const float b1 = __fmul_rn(0.5f, __fadd_rn(sample_depth(nr, fr, above, S), sample_depth(nr, fr, above + 1, S)));
    float denom = __fsub_rn(c1, c0);
    if (denom < 1e-5f) denom = 1.f;
    const float t = __fadd_rn(b0, __fmul_rn(__fdiv_rn(__fsub_rn(uu, c0), denom), __fsub_rn(b1, b0)));
    t_fine[(size_t)n * SF + j] = t;
    if (bins_out) bins_out[(size_t)n * SF + j] = inds;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Final ray march over the union of coarse and fine samples.  Entry e of a ray: e < S -> coarse sample i = e, else fine
// sample j = e - S.  The sort of unify_samples is realised as a rank (stable in e; equal depths are the same 3-D point and
// therefore carry identical density and colour, so the order among them does not change any output).
__global__ void __launch_bounds__(128) k_composite_merged(const float* __restrict__ dirs, const float* __restrict__ nearv,
                                                          const float* __restrict__ farv, int N, int S, int SF,
                                                          const FrameConst* __restrict__ fc,
                                                          const int* __restrict__ vid_c, const int* __restrict__ start_c,
                                                          const float* __restrict__ sigma_c, const float* __restrict__ rgb_c,
                                                          const float* __restrict__ noise_c,
                                                          const float* __restrict__ t_fine, const int* __restrict__ vid_f,
                                                          const int* __restrict__ start_f, const float* __restrict__ sigma_f,
                                                          const float* __restrict__ rgb_f, const float* __restrict__ noise_f,
                                                          int white_back, float* __restrict__ out_rgb, float* __restrict__ out_depth,
                                                          float* __restrict__ out_acc) {
  __shared__ float s_key[4][2 * kMaxS];     // depth of entry e; later alpha of entry e
  __shared__ float s_sorted[4][2 * kMaxS];  // depths in sorted order
  __shared__ float s_f[4][2 * kMaxS];       // (1 - alpha + 1e-10) in sorted order; later the exclusive transmittance
  __shared__ int s_pt[4][2 * kMaxS];        // compacted point index of entry e (fine: offset by 2^30), -1 when culled
  __shared__ unsigned short s_rank[4][2 * kMaxS];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int n = blockIdx.x * 4 + wl;
  if (n >= N) return;
  float* key = s_key[wl]; float* srt = s_sorted[wl]; float* fs = s_f[wl]; int* pt = s_pt[wl]; unsigned short* rank = s_rank[wl];
  const int M = S + SF;
  const float nr = nearv[n], fr = farv[n];
  const float dx = dirs[n * 3], dy = dirs[n * 3 + 1], dz = dirs[n * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  if (start_c[n] == start_c[n + 1] && start_f[n] == start_f[n + 1]) {
    // no surviving sample in either pass: every alpha is 0 -> exactly what the general path computes (0/0 -> NaN -> +inf -> clamp)
    if (lane == 0) {
      const float bg = white_back ? 1.f : -1.f;                                  // (0 + 1 - 0) * 2 - 1  |  0 * 2 - 1
      out_rgb[(size_t)n * 3] = bg; out_rgb[(size_t)n * 3 + 1] = bg; out_rgb[(size_t)n * 3 + 2] = bg;
      out_depth[n] = fminf(fmaxf(__int_as_float(0x7f800000), ordered_to_float(fc->dmin_bits)), ordered_to_float(fc->dmax_bits));
      out_acc[n] = 0.f;
    }
    return;
  }
  // phase 1: depths and compacted point indices (same ballot prefix as k_compact)
  int pos = start_c[n];
  for (int i0 = 0; i0 < S; i0 += 32) {
    const int i = i0 + lane;
    const int v = (i < S) ? vid_c[(size_t)n * S + i] : -1;
    const unsigned m = __ballot_sync(0xffffffffu, v >= 0);
    if (i < S) {
      key[i] = sample_depth(nr, fr, i, S);
      pt[i] = (v >= 0) ? pos + __popc(m & ((1u << lane) - 1u)) : -1;
    }
    pos += __popc(m);
  }
  pos = start_f[n];
  for (int j0 = 0; j0 < SF; j0 += 32) {
    const int j = j0 + lane;
    const int v = (j < SF) ? vid_f[(size_t)n * SF + j] : -1;
    const unsigned m = __ballot_sync(0xffffffffu, v >= 0);
    if (j < SF) {
      key[S + j] = t_fine[(size_t)n * SF + j];
      pt[S + j] = (v >= 0) ? ((pos + __popc(m & ((1u << lane) - 1u))) | (1 << 30)) : -1;
    }
    pos += __popc(m);
  }
  __syncwarp();
  // phase 2: rank of every entry under (depth, e).  The stratified coarse depths are monotone in i when far >= near, so a coarse
  // entry's rank among the coarse entries is i and a fine entry's is a binary search; only the fine list needs compare loops.
  if (fr >= nr) {
    for (int e = lane; e < M; e += 32) {
      const float t = key[e];
      int r;
      if (e < S) {
        r = e;
        for (int k = S; k < M; ++k) r += (key[k] < t) ? 1 : 0;                 // fine entries come after every coarse one in e
      } else {
        int lo = 0, hi = S;                                                    // #{i : t_i <= t}
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (key[mid] <= t) lo = mid + 1; else hi = mid; }
        r = lo;
        for (int k = S; k < M; ++k) { const float tk = key[k]; r += (tk < t || (tk == t && k < e)) ? 1 : 0; }
      }
      rank[e] = (unsigned short)r;
      srt[r] = t;
    }
  } else {
    for (int e = lane; e < M; e += 32) {
      const float t = key[e];
      int r = 0;
      for (int k = 0; k < M; ++k) {
        const float tk = key[k];
        r += (tk < t || (tk == t && k < e)) ? 1 : 0;
      }
      rank[e] = (unsigned short)r;
      srt[r] = t;
    }
  }
  __syncwarp();
  // phase 3: alpha per entry                                                                          ray_marcher.py:27-45
  for (int e = lane; e < M; e += 32) {
    const int r = rank[e];
    const float t = srt[r];
    const float delta = ((r == M - 1) ? 1e10f : (srt[r + 1] - t)) * dnorm;
    float alpha = 0.f;
    const int p = pt[e];
    if (p >= 0) {
      float sg;
      if (p & (1 << 30)) { sg = sigma_f[p & ~(1 << 30)]; if (noise_f) sg += noise_f[(size_t)n * SF + (e - S)]; }
      else { sg = sigma_c[p]; if (noise_c) sg += noise_c[p]; }
      alpha = 1.f - expf(-(fmaxf(sg, 0.f) * delta));
    }
    key[e] = alpha;
    fs[r] = 1.f - alpha + 1e-10f;
  }
  __syncwarp();
  // phase 4: exclusive product scan in sorted order                                                   ray_marcher.py:47-48
  float T = 1.f;
  for (int k0 = 0; k0 < M; k0 += 32) {
    const int k = k0 + lane;
    const float f = (k < M) ? fs[k] : 1.f;
    float incl = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= up;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    if (k < M) fs[k] = T * excl;
    T *= __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  // phase 5: weights and sums
  float cr = 0.f, cg = 0.f, cb = 0.f, wsum = 0.f, wdepth = 0.f;
  for (int e = lane; e < M; e += 32) {
    const int p = pt[e];
    if (p < 0) continue;
    const int r = rank[e];
    const float w = key[e] * fs[r];
    const float* c3 = (p & (1 << 30)) ? rgb_f + (size_t)(p & ~(1 << 30)) * 3 : rgb_c + (size_t)p * 3;
    cr += w * c3[0]; cg += w * c3[1]; cb += w * c3[2]; wsum += w; wdepth += w * srt[r];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cr += __shfl_xor_sync(0xffffffffu, cr, o);
    cg += __shfl_xor_sync(0xffffffffu, cg, o);
    cb += __shfl_xor_sync(0xffffffffu, cb, o);
    wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    wdepth += __shfl_xor_sync(0xffffffffu, wdepth, o);
  }
  if (lane == 0) {
    float depth = wdepth / wsum;                                            // ray_marcher.py:53-57 (fine depths lie inside the
    if (depth != depth) depth = __int_as_float(0x7f800000);                 // coarse range, so min/max(all_depths) is unchanged)
    depth = fminf(fmaxf(depth, ordered_to_float(fc->dmin_bits)), ordered_to_float(fc->dmax_bits));
    if (white_back) { cr = cr + 1.f - wsum; cg = cg + 1.f - wsum; cb = cb + 1.f - wsum; }
    out_rgb[(size_t)n * 3] = cr * 2.f - 1.f;
    out_rgb[(size_t)n * 3 + 1] = cg * 2.f - 1.f;
    out_rgb[(size_t)n * 3 + 2] = cb * 2.f - 1.f;
    out_depth[n] = depth;
    out_acc[n] = wsum;
  }
}

// dense per-sample taps of the fine pass (debug only): sigma = -80 / rgb = 0 where culled                renderer.py:364-371
__global__ void k_fill_dense(float* __restrict__ sigma, float* __restrict__ rgb, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    if (sigma) sigma[i] = -80.f;
    if (rgb) { rgb[i * 3] = 0.f; rgb[i * 3 + 1] = 0.f; rgb[i * 3 + 2] = 0.f; }
  }
}
__global__ void k_scatter_dense(const int* __restrict__ point_sample, const float* __restrict__ sg, const float* __restrict__ c3,
                                int64_t P, float* __restrict__ sigma, float* __restrict__ rgb) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) {
    const int s = point_sample[p];
    if (sigma) sigma[s] = sg[p];
    if (rgb) { rgb[(size_t)s * 3] = c3[p * 3]; rgb[(size_t)s * 3 + 1] = c3[p * 3 + 1]; rgb[(size_t)s * 3 + 2] = c3[p * 3 + 2]; }
  }
}

int run_importance_sample(const SherfRays& rays, const int* ray_start, const int* point_sample, const float* sigma, const float* noise,
                          const float* w_in, const float* u, float* t_fine, int* bins_out, float* w_out, cudaStream_t st) {
  const int N = rays.n_rays, S = rays.n_samples, SF = rays.n_importance;
  if (w_in)
    k_importance_sample<false><<<ceil_div(N, 4), 128, 0, st>>>(rays.dirs, rays.near_, rays.far_, N, S, SF, nullptr, nullptr, nullptr, nullptr,
                                                               w_in, u, t_fine, bins_out, w_out);
  else
    k_importance_sample<true><<<ceil_div(N, 4), 128, 0, st>>>(rays.dirs, rays.near_, rays.far_, N, S, SF, ray_start, point_sample, sigma,
                                                              noise, nullptr, u, t_fine, bins_out, w_out);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

int run_composite_merged(const SherfRays& rays, const FrameConst* fc, const int* vid_c, const int* start_c, const float* sigma_c,
                         const float* rgb_c, const float* noise_c, const float* t_fine, const int* vid_f, const int* start_f,
                         const float* sigma_f, const float* rgb_f, const float* noise_f, int white_back, const SherfOut& out,
                         cudaStream_t st) {
  k_composite_merged<<<ceil_div(rays.n_rays, 4), 128, 0, st>>>(rays.dirs, rays.near_, rays.far_, rays.n_rays, rays.n_samples,
                                                               rays.n_importance, fc, vid_c, start_c, sigma_c, rgb_c, noise_c, t_fine,
                                                               vid_f, start_f, sigma_f, rgb_f, noise_f, white_back, out.rgb, out.depth,
                                                               out.acc);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

int run_dense_taps(const int* point_sample, const float* sg, const float* c3, int64_t P, int64_t n_dense, float* sigma, float* rgb,
                   cudaStream_t st) {
  if (!sigma && !rgb) return SHERF_OK;
  k_fill_dense<<<ceil_div(n_dense, 256), 256, 0, st>>>(sigma, rgb, n_dense);
  SHERF_LAUNCH_CHECK();
  if (P > 0) {
    k_scatter_dense<<<ceil_div(P, 256), 256, 0, st>>>(point_sample, sg, c3, P, sigma, rgb);
    SHERF_LAUNCH_CHECK();
  }
  return SHERF_OK;
}

}  // namespace sherf
