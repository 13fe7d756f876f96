// Stage 1: sample points along rays, exact nearest posed-vertex search within the 5 cm cull radius
// (uniform grid, 27-cell neighbourhood), ordered compaction of the survivors.
// Replaces renderer.py:299-321 (sample_stratified, SMPL-space transform, knn_points #1, mask, boolean-index
// compaction).  Index bookkeeping is bit-exact against oracle/port.py (see its header for the rounding order).
#include "common.cuh"

namespace sherf {

// Search the 27-cell neighbourhood of q's cell for the nearest vertex; returns best squared distance and id
// under the lexicographic (d2, id) order = "smallest index wins ties".
__device__ __forceinline__ void nn_27(const GridDesc& g, const int* __restrict__ cell_start, const float4* __restrict__ gv,
                                      float qx, float qy, float qz, int cx, int cy, int cz, float& best, int& best_id) {
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
  for (int z = max(cz - 1, 0); z <= min(cz + 1, g.dim[2] - 1); ++z) {
    for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
      const int row = (z * g.dim[1] + y) * g.dim[0];
      const int b = cell_start[row + x0], e = cell_start[row + x1 + 1];
      for (int k = b; k < e; ++k) {
        const float4 v = gv[k];
        const float d2 = dist2_xyz(qx, qy, qz, v.x, v.y, v.z);
        const int id = __float_as_int(v.w);
        if (d2 < best || (d2 == best && id < best_id)) { best = d2; best_id = id; }
      }
    }
  }
}

// One warp per ray; lane handles samples lane, lane+32, ...
// Writes sample_vid[n*S+i] (vertex id, or -1 when culled) and ray_count[n].
__global__ void __launch_bounds__(256) k_cull(const float* __restrict__ origins, const float* __restrict__ dirs,
                                              const float* __restrict__ nearv, const float* __restrict__ farv, int N, int S,
                                              const FrameConst* __restrict__ fcp, const int* __restrict__ cell_start,
                                              const float4* __restrict__ gv, const unsigned char* __restrict__ occ, float thr,
                                              const float* __restrict__ depths, int* __restrict__ sample_vid, int* __restrict__ ray_count) {
  __shared__ FrameConst fc;
  for (int i = threadIdx.x; i < (int)(sizeof(FrameConst) / 4); i += blockDim.x) ((int*)&fc)[i] = ((const int*)fcp)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float ox = origins[n * 3], oy = origins[n * 3 + 1], oz = origins[n * 3 + 2];
  const float dx = dirs[n * 3], dy = dirs[n * 3 + 1], dz = dirs[n * 3 + 2];
  const float nr = nearv[n], fr = farv[n];
  const GridDesc& g = fc.g1;
  int count = 0;
  for (int i0 = 0; i0 < S; i0 += 32) {
    const int i = i0 + lane;
    int vid = -1;
    if (i < S) {
      const float t = depths ? depths[(size_t)n * S + i] : sample_depth(nr, fr, i, S);   // fine pass: importance-sampled depths
      float p[3], q[3];
      p[0] = __fsub_rn(mul_add_sep(t, dx, ox), fc.Th_tgt[0]);
      p[1] = __fsub_rn(mul_add_sep(t, dy, oy), fc.Th_tgt[1]);
      p[2] = __fsub_rn(mul_add_sep(t, dz, oz), fc.Th_tgt[2]);
      rowvec_mat3(p, fc.R_tgt, q);
      const int cx = grid_coord(q[0], g.origin[0], g.inv_cell, g.dim[0]);
      const int cy = grid_coord(q[1], g.origin[1], g.inv_cell, g.dim[1]);
      const int cz = grid_coord(q[2], g.origin[2], g.inv_cell, g.dim[2]);
      if (cx >= 0 && cx < g.dim[0] && cy >= 0 && cy < g.dim[1] && cz >= 0 && cz < g.dim[2] &&
          occ[(cz * g.dim[1] + cy) * g.dim[0] + cx]) {
        float best = 3.0e38f;
        int bid = 0x7fffffff;
        nn_27(g, cell_start, gv, q[0], q[1], q[2], cx, cy, cz, best, bid);
        if (best < thr) vid = bid;
      }
      sample_vid[(size_t)n * S + i] = vid;
    }
    count += __popc(__ballot_sync(0xffffffffu, vid >= 0));
  }
  if (lane == 0) ray_count[n] = count;
}

// Exclusive scan of ray_count[0..N) -> ray_start[0..N] in two coalesced passes over 1024-ray blocks.
__global__ void __launch_bounds__(1024) k_ray_block_sums(const int* __restrict__ cnt, int N, int* __restrict__ bsum) {
  __shared__ int red[32];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  int v = i < N ? cnt[i] : 0;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    int x = red[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (threadIdx.x == 0) bsum[blockIdx.x] = x;
  }
}

__global__ void __launch_bounds__(1024) k_scan_rays(const int* __restrict__ cnt, const int* __restrict__ bsum, int nblocks, int N,
                                                    int* __restrict__ start, int64_t* total_out) {
  __shared__ int red[32];
  __shared__ int warp_off[32];
  __shared__ int block_off;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // offset of this block = sum of the preceding block sums (nblocks is small: N / 1024)
  int pre = 0;
  for (int b = tid; b < (int)blockIdx.x; b += 1024) pre += bsum[b];
  for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
  if (lane == 0) red[w] = pre;
  __syncthreads();
  if (tid < 32) {
    int x = red[tid];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (tid == 0) block_off = x;
  }
  __syncthreads();
  const int i = blockIdx.x * 1024 + tid;
  const int v = i < N ? cnt[i] : 0;
  int incl = v;
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  __syncthreads();
  if (lane == 31) red[w] = incl;
  __syncthreads();
  if (tid < 32) {
    const int x = red[tid];
    int y = x;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, y, o); if (tid >= o) y += t; }
    warp_off[tid] = y - x;
  }
  __syncthreads();
  const int excl = block_off + warp_off[w] + incl - v;
  if (i < N) start[i] = excl;
  if (i == N - 1) { start[N] = excl + v; *total_out = excl + v; }
}

// Ordered scatter: point_sample[pos] = n*S+i, point_vid[pos] = vid, pos ascending in row-major sample order.
__global__ void __launch_bounds__(256) k_compact(const int* __restrict__ sample_vid, const int* __restrict__ ray_start, int N, int S,
                                                 int* __restrict__ point_sample, int* __restrict__ point_vid) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  int pos = ray_start[n];
  if (ray_start[n + 1] == pos) return;
  for (int i0 = 0; i0 < S; i0 += 32) {
    const int i = i0 + lane;
    const int vid = (i < S) ? sample_vid[(size_t)n * S + i] : -1;
    const unsigned m = __ballot_sync(0xffffffffu, vid >= 0);
    if (vid >= 0) {
      const int p = pos + __popc(m & ((1u << lane) - 1u));
      point_sample[p] = n * S + i;
      point_vid[p] = vid;
    }
    pos += __popc(m);
  }
}

int run_cull(const SherfRays& rays, int S, const float* depths, const FrameTables& ft, int* sample_vid, int* ray_count, int* block_sums,
             int* ray_start, int64_t* total_dev, int* point_sample, int* point_vid, cudaStream_t st) {
  const int N = rays.n_rays;
  const float thr = (float)(0.05 * 0.05);        // `distance < 0.05 ** 2` compares in fp32 (renderer.py:318-319)
  k_cull<<<ceil_div(N, 8), 256, 0, st>>>(rays.origins, rays.dirs, rays.near_, rays.far_, N, S, ft.fc, ft.g1_cell_start, ft.g1_verts,
                                         ft.g1_occ, thr, depths, sample_vid, ray_count);
  SHERF_LAUNCH_CHECK();
  const int nb = ceil_div(N, 1024);
  k_ray_block_sums<<<nb, 1024, 0, st>>>(ray_count, N, block_sums);
  SHERF_LAUNCH_CHECK();
  k_scan_rays<<<nb, 1024, 0, st>>>(ray_count, block_sums, nb, N, ray_start, total_dev);
  SHERF_LAUNCH_CHECK();
  k_compact<<<ceil_div(N, 8), 256, 0, st>>>(sample_vid, ray_start, N, S, point_sample, point_vid);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

}  // namespace sherf
