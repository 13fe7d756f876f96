// Sparse 3-D encoder: the reference's SparseConvNet (renderer.py:708-797, layers :814-871) evaluated on the
// sparse voxel tensor triplane.py:137 builds from the SMPL vertices, producing the three densified pyramid levels the render path
// samples (renderer.py:762,771,780).  spconv is replaced by dense INDEX grids (voxel -> row, -1 = inactive; 54 MB for the
// 96x320x384 canonical box) instead of hash tables: <= 6 890 seed voxels make every level tiny (<= ~60 k active rows), so the convs
// are gather-form warp-per-output-row kernels with weights re-laid-out as [offset][c_in][c_out] (coalesced across lanes); the whole
// encoder is a few GFLOP and latency-bound.  Semantics and the duplicate-voxel convention: oracle/sparse_encoder.py (header).
#include "common.cuh"
#include "stages.cuh"

namespace sherf {

struct SpDims { int d[3]; };
__device__ __forceinline__ int sp_cell(const SpDims& s, int z, int y, int x) { return (z * s.d[1] + y) * s.d[2] + x; }
__device__ __forceinline__ bool sp_inside(const SpDims& s, int z, int y, int x) {
  return z >= 0 && z < s.d[0] && y >= 0 && y < s.d[1] && x >= 0 && x < s.d[2];
}

// ---- level 0: vertices -> unique voxels (smallest vertex index represents a voxel) -> rows ----
__global__ void k_sp_claim(const int* __restrict__ coord, int n, SpDims s, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = coord[i * 3], y = coord[i * 3 + 1], x = coord[i * 3 + 2];
  if (sp_inside(s, z, y, x)) atomicMin(reinterpret_cast<unsigned*>(&idx[sp_cell(s, z, y, x)]), (unsigned)i);   // -1 = 0xffffffff is the largest
}
__global__ void k_sp_rows0(const int* __restrict__ coord, const float* __restrict__ feat, int n, int C, SpDims s, const int* __restrict__ idx,
                           int* __restrict__ count, int* __restrict__ rowof, int* __restrict__ coords_out, float* __restrict__ F) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = coord[i * 3], y = coord[i * 3 + 1], x = coord[i * 3 + 2];
  int row = -1;
  if (sp_inside(s, z, y, x) && idx[sp_cell(s, z, y, x)] == i) {
    row = atomicAdd(count, 1);
    coords_out[row * 3] = z; coords_out[row * 3 + 1] = y; coords_out[row * 3 + 2] = x;
    for (int c = 0; c < C; ++c) F[(size_t)row * C + c] = feat[(size_t)i * C + c];
  }
  rowof[i] = row;
}
__global__ void k_sp_index0(const int* __restrict__ coord, int n, SpDims s, const int* __restrict__ rowof, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || rowof[i] < 0) return;
  idx[sp_cell(s, coord[i * 3], coord[i * 3 + 1], coord[i * 3 + 2])] = rowof[i];
}

// ---- weights [c_out][27][c_in] (spconv KRSC) -> [27][c_in][c_out]; BatchNorm (eval) folded into scale / shift ----
__global__ void k_sp_pack(const float* __restrict__ W, const float* __restrict__ bw, const float* __restrict__ bb, const float* __restrict__ bm,
                          const float* __restrict__ bv, int cin, int cout, float* __restrict__ Wt, float* __restrict__ scale_shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 27 * cin * cout) {
    const int co = i % cout, t = i / cout, ci = t % cin, o = t / cin;
    Wt[i] = W[((size_t)co * 27 + o) * cin + ci];
  }
  if (i < cout) {
    const float sc = bw[i] / sqrtf(bv[i] + 1e-3f);                 // BatchNorm1d(eps=1e-3), renderer.py:822
    scale_shift[i] = sc;
    scale_shift[cout + i] = bb[i] - bm[i] * sc;
  }
}

// ---- gather-form convolution, one warp per output row.  DOWN = false: SubMConv3d (in = out sites, taps p + d);
//      DOWN = true: SparseConv3d k3 s2 p1 (taps 2 o - 1 + k on the finer level) ----
template <bool DOWN>
__global__ void __launch_bounds__(256) k_sp_conv(const int* __restrict__ coords, const int* __restrict__ count, SpDims sin,
                                                 const int* __restrict__ idx_in, const float* __restrict__ Fin, int cin, int cout,
                                                 const float* __restrict__ Wt, const float* __restrict__ scale_shift, float* __restrict__ Fout) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= *count) return;
  const int z = coords[r * 3], y = coords[r * 3 + 1], x = coords[r * 3 + 2];
  float acc[3] = {0.f, 0.f, 0.f};                                   // c_out <= 96: channels lane, lane + 32, lane + 64
  for (int o = 0; o < 27; ++o) {
    const int kz = o / 9, ky = (o / 3) % 3, kx = o % 3;
    const int pz = DOWN ? 2 * z - 1 + kz : z + kz - 1, py = DOWN ? 2 * y - 1 + ky : y + ky - 1, px = DOWN ? 2 * x - 1 + kx : x + kx - 1;
    if (!sp_inside(sin, pz, py, px)) continue;
    const int j = idx_in[sp_cell(sin, pz, py, px)];
    if (j < 0) continue;
    const float* a = Fin + (size_t)j * cin;
    const float* w = Wt + (size_t)o * cin * cout;
    for (int ci = 0; ci < cin; ++ci) {
      const float av = a[ci];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int co = lane + 32 * t;
        if (co < cout) acc[t] = fmaf(av, w[(size_t)ci * cout + co], acc[t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int co = lane + 32 * t;
    if (co < cout) Fout[(size_t)r * cout + co] = fmaxf(acc[t] * scale_shift[co] + scale_shift[cout + co], 0.f);     // BN + ReLU
  }
}

// ---- output sites of a strided conv: o is active iff some active input p = 2 o - 1 + k ----
__global__ void k_sp_down_sites(const int* __restrict__ coords_in, const int* __restrict__ count_in, SpDims sout, int* __restrict__ idx_out,
                                int* __restrict__ count_out, int* __restrict__ coords_out, int cap_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = t / 27, o = t % 27;
  if (r >= *count_in) return;
  const int kz = o / 9, ky = (o / 3) % 3, kx = o % 3;
  const int oz2 = coords_in[r * 3] + 1 - kz, oy2 = coords_in[r * 3 + 1] + 1 - ky, ox2 = coords_in[r * 3 + 2] + 1 - kx;
  if ((oz2 | oy2 | ox2) & 1) return;
  const int oz = oz2 >> 1, oy = oy2 >> 1, ox = ox2 >> 1;
  if (oz2 < 0 || oy2 < 0 || ox2 < 0 || !sp_inside(sout, oz, oy, ox)) return;
  const int cell = sp_cell(sout, oz, oy, ox);
  if (atomicCAS(&idx_out[cell], -1, -2) == -1) {
    const int row = atomicAdd(count_out, 1);
    if (row < cap_out) {
      coords_out[row * 3] = oz; coords_out[row * 3 + 1] = oy; coords_out[row * 3 + 2] = ox;
      idx_out[cell] = row;
    }
  }
}

// ---- .dense(): [C][D][H][W], zero-filled by the caller ----
__global__ void k_sp_densify(const int* __restrict__ coords, const int* __restrict__ count, SpDims s, const float* __restrict__ F, int C,
                             float* __restrict__ vol) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(t / C), c = (int)(t % C);
  if (r >= *count) return;
  vol[(size_t)c * s.d[0] * s.d[1] * s.d[2] + sp_cell(s, coords[r * 3], coords[r * 3 + 1], coords[r * 3 + 2])] = F[(size_t)r * C + c];
}

// level l = 0..3: dims, row capacity
struct SpPlan { SpDims dims[4]; int cap[4]; size_t cells[4]; };
static void sp_plan(int n, const int32_t* out_sh, SpPlan& p) {
  for (int a = 0; a < 3; ++a) p.dims[0].d[a] = out_sh[a];
  for (int l = 1; l < 4; ++l) for (int a = 0; a < 3; ++a) p.dims[l].d[a] = (p.dims[l - 1].d[a] + 2 - 3) / 2 + 1;
  int64_t grow = n;
  for (int l = 0; l < 4; ++l) {
    p.cells[l] = (size_t)p.dims[l].d[0] * p.dims[l].d[1] * p.dims[l].d[2];
    if (l > 0) grow *= 8;                                            // a strided conv creates at most 8 output sites per input site
    p.cap[l] = (int)((int64_t)p.cells[l] < grow ? (int64_t)p.cells[l] : grow);
  }
}

struct SpScratch { int* idx[4]; int* coords[4]; int* count; int* rowof; float* F[2]; float* Wt; float* ss; };
static size_t sp_carve(char* base, int n, const SpPlan& p, SpScratch& s) {
  size_t off = 0;
  auto take = [&](size_t bytes) { off = (off + 255) & ~(size_t)255; char* q = base ? base + off : nullptr; off += bytes; return q; };
  int maxcap = 0;
  for (int l = 0; l < 4; ++l) {
    s.idx[l] = (int*)take(p.cells[l] * sizeof(int));
    s.coords[l] = (int*)take((size_t)p.cap[l] * 3 * sizeof(int));
    if (p.cap[l] > maxcap) maxcap = p.cap[l];
  }
  s.count = (int*)take(4 * sizeof(int));
  s.rowof = (int*)take((size_t)n * sizeof(int));
  s.F[0] = (float*)take((size_t)maxcap * 96 * sizeof(float));
  s.F[1] = (float*)take((size_t)maxcap * 96 * sizeof(float));
  s.Wt = (float*)take((size_t)27 * 96 * 96 * sizeof(float));
  s.ss = (float*)take(2 * 96 * sizeof(float));
  return off + 256;
}

size_t sparse_encoder_scratch_bytes(int n, const int32_t* out_sh) {
  SpPlan p; sp_plan(n, out_sh, p);
  SpScratch s;
  return sp_carve(nullptr, n, p, s) + 256;
}

int run_sparse_encode(const SherfSparseEncoder& enc, const int* coord, const float* feat, int n, const int32_t* out_sh, float* const* vols,
                      void* scratch, size_t scratch_bytes, cudaStream_t st) {
  SpPlan p; sp_plan(n, out_sh, p);
  char* base = (char*)scratch;
  const size_t mis = ((size_t)base) & 255;
  if (mis) base += 256 - mis;
  SpScratch s;
  const size_t need = sp_carve(base, n, p, s);
  if (need + 256 > scratch_bytes) { set_error("sparse-encoder scratch too small: need %zu bytes, have %zu", need + 256, scratch_bytes); return SHERF_E_SCRATCH; }
  for (int l = 0; l < 4; ++l) SHERF_CUDA_OK(cudaMemsetAsync(s.idx[l], 0xff, p.cells[l] * sizeof(int), st));
  SHERF_CUDA_OK(cudaMemsetAsync(s.count, 0, 4 * sizeof(int), st));
  const int C0 = enc.conv[0].c_in;
  k_sp_claim<<<ceil_div(n, 256), 256, 0, st>>>(coord, n, p.dims[0], s.idx[0]);
  SHERF_LAUNCH_CHECK();
  k_sp_rows0<<<ceil_div(n, 256), 256, 0, st>>>(coord, feat, n, C0, p.dims[0], s.idx[0], s.count, s.rowof, s.coords[0], s.F[0]);
  SHERF_LAUNCH_CHECK();
  k_sp_index0<<<ceil_div(n, 256), 256, 0, st>>>(coord, n, p.dims[0], s.rowof, s.idx[0]);
  SHERF_LAUNCH_CHECK();
  int level = 0, cur = 0, emitted = 0;
  // execution order and the levels emitted after conv1 / conv2 / conv3: renderer.py:756-782 (num_layers = 4)
  static const int emit_after[SHERF_SPARSE_CONVS] = {0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int c = 0; c < SHERF_SPARSE_CONVS; ++c) {
    const SherfSparseConv& L = enc.conv[c];
    if (L.c_in > 96 || L.c_out > 96 || L.c_in <= 0 || L.c_out <= 0) { set_error("sparse conv %d: unsupported channels %d -> %d", c, L.c_in, L.c_out); return SHERF_E_UNSUPPORTED; }
    k_sp_pack<<<ceil_div(27 * L.c_in * L.c_out, 256), 256, 0, st>>>(L.weight, L.bn_weight, L.bn_bias, L.bn_mean, L.bn_var, L.c_in, L.c_out, s.Wt, s.ss);
    SHERF_LAUNCH_CHECK();
    if (L.kind == 0) {
      k_sp_conv<false><<<ceil_div(p.cap[level], 8), 256, 0, st>>>(s.coords[level], s.count + level, p.dims[level], s.idx[level], s.F[cur], L.c_in,
                                                                  L.c_out, s.Wt, s.ss, s.F[cur ^ 1]);
      SHERF_LAUNCH_CHECK();
    } else {
      if (level >= 3) { set_error("sparse encoder: too many strided convs"); return SHERF_E_INVALID; }
      k_sp_down_sites<<<ceil_div((int64_t)p.cap[level] * 27, 256), 256, 0, st>>>(s.coords[level], s.count + level, p.dims[level + 1], s.idx[level + 1],
                                                                               s.count + level + 1, s.coords[level + 1], p.cap[level + 1]);
      SHERF_LAUNCH_CHECK();
      k_sp_conv<true><<<ceil_div(p.cap[level + 1], 8), 256, 0, st>>>(s.coords[level + 1], s.count + level + 1, p.dims[level], s.idx[level], s.F[cur],
                                                                     L.c_in, L.c_out, s.Wt, s.ss, s.F[cur ^ 1]);
      SHERF_LAUNCH_CHECK();
      ++level;
    }
    cur ^= 1;
    if (emit_after[c]) {
      if (emitted >= 3 || level != emitted + 1) { set_error("sparse encoder: unexpected layer order"); return SHERF_E_INVALID; }
      SHERF_CUDA_OK(cudaMemsetAsync(vols[emitted], 0, p.cells[level] * (size_t)L.c_out * sizeof(float), st));
      k_sp_densify<<<ceil_div((int64_t)p.cap[level] * L.c_out, 256), 256, 0, st>>>(s.coords[level], s.count + level, p.dims[level], s.F[cur], L.c_out,
                                                                                 vols[emitted]);
      SHERF_LAUNCH_CHECK();
      ++emitted;
    }
  }
  return SHERF_OK;
}

}  // namespace sherf
