// Dataset-side ray setup on the GPU: get_rays (RenderPeople_dataset.py:14-27) and get_near_far
// (:68-101) with the near/far fill of sample_ray_RenderPeople_batch (:129-134), one thread per pixel.  The dataset
// computes these in numpy float64 and casts to float32 at the end; the kernel does the same in fp64 (262 144 rays x ~150
// flops: negligible even at B200's fp64 rate) so that results agree to the last float32 bit except where a float64 sum
// order differs from BLAS.  Removes the per-frame host ray generation and the 3 x N float H2D copy of streamed sequences.
#include "common.cuh"
#include "stages.cuh"

namespace sherf {

struct RayCam { double Kinv[9]; double R[9]; double T[3]; double o[3]; double bmin[3]; double bmax[3]; };

__global__ void __launch_bounds__(256) k_generate_rays(const RayCam c, int H, int W, float* __restrict__ origins, float* __restrict__ dirs,
                                                       float* __restrict__ nearv, float* __restrict__ farv,
                                                       unsigned char* __restrict__ mask_at_box) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= H * W) return;
  const double px = (double)(float)(n % W), py = (double)(float)(n / W);      // np.arange(.., dtype=float32), indexing='xy'
  // pixel_camera = [i, j, 1] @ inv(K).T ; pixel_world = (pixel_camera - T) @ R ; rays_d = pixel_world - rays_o       :20-25
  double pc[3], pw[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) pc[k] = (px * c.Kinv[3 * k] + py * c.Kinv[3 * k + 1]) + c.Kinv[3 * k + 2];
#pragma unroll
  for (int k = 0; k < 3; ++k) pc[k] -= c.T[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) pw[k] = (pc[0] * c.R[k] + pc[1] * c.R[3 + k]) + pc[2] * c.R[6 + k];
  float o[3], d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { o[k] = (float)c.o[k]; d[k] = (float)(pw[k] - c.o[k]); }   // .astype(np.float32)           :125-126
  // get_near_far                                                                                                     :68-101
#pragma unroll
  for (int k = 0; k < 3; ++k) if (d[k] == 0.0f) d[k] = 1e-8f;                              // ray_d[ray_d==0.0] = 1e-8 (in place)
  const double eps = 1e-6;
  int count = 0;
  double dist[2] = {0.0, 0.0};
#pragma unroll
  for (int f = 0; f < 6; ++f) {                      // plane order of bounds.ravel(): min_x, min_y, min_z, max_x, max_y, max_z
    const int ax = f % 3;
    const double plane = f < 3 ? c.bmin[ax] : c.bmax[ax];
    const double t = (plane - (double)o[ax]) / (double)d[ax];
    double p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = t * (double)d[k] + (double)o[k];
    const bool in = p[0] >= c.bmin[0] - eps && p[0] <= c.bmax[0] + eps && p[1] >= c.bmin[1] - eps && p[1] <= c.bmax[1] + eps &&
                    p[2] >= c.bmin[2] - eps && p[2] <= c.bmax[2] + eps;
    if (in) {
      if (count < 2) {
        const double ex = p[0] - (double)o[0], ey = p[1] - (double)o[1], ez = p[2] - (double)o[2];
        dist[count] = sqrt((ex * ex + ey * ey) + ez * ez);                                  // np.linalg.norm(float64)
      }
      ++count;
    }
  }
  const bool hit = count == 2;                                                              // exactly two intersections  :88
  float nr = 0.f, fr = 1.f;                                                                 // near_all / far_all fill    :129-134
  if (hit) {
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));   // norm(float32)
    const double d0 = dist[0] / (double)nrm, d1 = dist[1] / (double)nrm;
    nr = (float)fmin(d0, d1);
    fr = (float)fmax(d0, d1);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) { origins[(size_t)n * 3 + k] = o[k]; dirs[(size_t)n * 3 + k] = d[k]; }
  nearv[n] = nr;
  farv[n] = fr;
  if (mask_at_box) mask_at_box[n] = hit ? 1 : 0;
}

static void inv3(const double* m, double* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  const double r = 1.0 / det;
  o[0] = (e * i - f * h) * r; o[1] = (c * h - b * i) * r; o[2] = (b * f - c * e) * r;
  o[3] = (f * g - d * i) * r; o[4] = (a * i - c * g) * r; o[5] = (c * d - a * f) * r;
  o[6] = (d * h - e * g) * r; o[7] = (b * g - a * h) * r; o[8] = (a * e - b * d) * r;
}

int run_generate_rays(const double* K, const double* R, const double* T, int H, int W, const double* bounds, float* origins, float* dirs,
                      float* nearv, float* farv, unsigned char* mask_at_box, cudaStream_t st) {
  RayCam c;
  inv3(K, c.Kinv);
  for (int k = 0; k < 9; ++k) c.R[k] = R[k];
  for (int k = 0; k < 3; ++k) {
    c.T[k] = T[k];
    c.o[k] = -(R[k] * T[0] + R[3 + k] * T[1] + R[6 + k] * T[2]);                            // rays_o = -R^T T              :16
    c.bmin[k] = bounds[k] - 0.01;                                                          // bounds + [-0.01, 0.01]       :70
    c.bmax[k] = bounds[3 + k] + 0.01;
  }
  k_generate_rays<<<ceil_div((int64_t)H * W, 256), 256, 0, st>>>(c, H, W, origins, dirs, nearv, farv, mask_at_box);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

}  // namespace sherf
