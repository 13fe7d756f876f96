// Stage 4: front-to-back alpha compositing per ray over the surviving samples (culled samples carry sigma = -80 ->
// relu -> alpha = 0 and contribute a transmittance factor of exactly fl(1 + 1e-10) = 1).  Replaces the scatter-back of
// renderer.py:364-371 and MipRayMarcher2.run_forward (ray_marcher.py:25-64, clamp_mode 'relu').  One warp per ray.
#include "common.cuh"
#include "stages.cuh"

namespace sherf {

__global__ void __launch_bounds__(256) k_composite(const float* __restrict__ dirs, const float* __restrict__ nearv,
                                                   const float* __restrict__ farv, int N, int S, const FrameConst* __restrict__ fc,
                                                   const int* __restrict__ ray_start, const int* __restrict__ point_sample,
                                                   const float* __restrict__ sigma, const float* __restrict__ rgb,
                                                   const float* __restrict__ noise, int white_back, float* __restrict__ out_rgb,
                                                   float* __restrict__ out_depth, float* __restrict__ out_acc) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int b = ray_start[n], e = ray_start[n + 1];
  const float dx = dirs[n * 3], dy = dirs[n * 3 + 1], dz = dirs[n * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);                     // ray_marcher.py:29
  const float nr = nearv[n], fr = farv[n];
  float T = 1.f;                                                             // transmittance carried across 32-point passes
  float cr = 0.f, cg = 0.f, cb = 0.f, wsum = 0.f, wdepth = 0.f;
  for (int base = b; base < e; base += 32) {
    const int p = base + lane;
    float alpha = 0.f, t = 0.f, r = 0.f, g = 0.f, bl = 0.f;
    if (p < e) {
      const int s = point_sample[p];
      const int i = s - n * S;
      t = sample_depth(nr, fr, i, S);
      const float delta = ((i == S - 1) ? 1e10f : (sample_depth(nr, fr, i + 1, S) - t)) * dnorm;   // ray_marcher.py:27-29
      float sg = sigma[p];
      if (noise) sg += noise[p];                                            // per SURVIVING point, compacted order: renderer.py:435-436
      alpha = 1.f - expf(-(fmaxf(sg, 0.f) * delta));                        // ray_marcher.py:39-45
      r = rgb[(size_t)p * 3]; g = rgb[(size_t)p * 3 + 1]; bl = rgb[(size_t)p * 3 + 2];
    }
    // exclusive product scan of (1 - alpha + 1e-10) across the warp        ray_marcher.py:47-48
    const float f = (p < e) ? (1.f - alpha + 1e-10f) : 1.f;
    float incl = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= up;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    const float w = alpha * (T * excl);
    cr += w * r; cg += w * g; cb += w * bl; wsum += w; wdepth += w * t;
    T *= __shfl_sync(0xffffffffu, incl, 31);
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
    float depth = wdepth / wsum;                                            // 0/0 -> NaN -> +inf -> clamp   ray_marcher.py:53-57
    if (depth != depth) depth = __int_as_float(0x7f800000);
    depth = fminf(fmaxf(depth, ordered_to_float(fc->dmin_bits)), ordered_to_float(fc->dmax_bits));
    if (white_back) { cr = cr + 1.f - wsum; cg = cg + 1.f - wsum; cb = cb + 1.f - wsum; }
    out_rgb[(size_t)n * 3] = cr * 2.f - 1.f;
    out_rgb[(size_t)n * 3 + 1] = cg * 2.f - 1.f;
    out_rgb[(size_t)n * 3 + 2] = cb * 2.f - 1.f;
    out_depth[n] = depth;
    out_acc[n] = wsum;
  }
}

int run_composite(const SherfRays& rays, const FrameConst* fc, const int* ray_start, const int* point_sample, const float* sigma,
                  const float* rgb, const float* noise, int white_back, const SherfOut& out, cudaStream_t st) {
  k_composite<<<ceil_div(rays.n_rays, 8), 256, 0, st>>>(rays.dirs, rays.near_, rays.far_, rays.n_rays, rays.n_samples, fc, ray_start,
                                                        point_sample, sigma, rgb, noise, white_back, out.rgb, out.depth, out.acc);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

}  // namespace sherf
