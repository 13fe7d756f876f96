// Sparse 3-D encoder: the reference's SparseConvNet (renderer.py:708-797, layers :814-871) evaluated on the
// sparse voxel tensor triplane.py:137 builds from the SMPL vertices, producing the three densified pyramid levels the render path
// samples (renderer.py:762,771,780).  spconv is replaced by dense INDEX grids (voxel -> row, -1 = inactive; 54 MB for the
// 96x320x384 canonical box) instead of hash tables: <= 6 890 seed voxels make every level tiny (<= ~60 k active rows), so the convs
// are gather-form warp-per-output-row kernels with weights re-laid-out as [offset][c_in][c_out] (coalesced across lanes); the whole
// encoder is a few GFLOP and latency-bound.  Semantics and the duplicate-voxel convention: oracle/sparse_encoder.py (header).
#include "common.cuh"
#include "stages.cuh"
#include <cstdio>
#include <cstdlib>

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
    if (co < cout) Fout[(size_t)r * cout + co] = scale_shift ? fmaxf(acc[t] * scale_shift[co] + scale_shift[cout + co], 0.f) : acc[t];   // BN + ReLU | raw
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

// ---- the convolutions on the tensor cores: every layer is a linear layer with K = 27 c_in whose A rows are gathered through a neighbour
//      table (mlp_umma.cu: launch_umma_spconv, 3xTF32 split products, split-K partial tiles summed in order) ----
// MODE 0: SubMConv3d, table of the output (= input) sites of a level: slot o -> row at p + (k - 1)
// MODE 1: strided convolution, table of the OUTPUT sites (coarse level): slot o -> fine row at 2 q - 1 + k
// MODE 2: strided convolution, table of the INPUT sites (fine level): slot o -> coarse row q with 2 q - 1 + k = p (its gradient reaches p through W[:, o, :])
template <int MODE>
__global__ void k_sp_table(const int* __restrict__ coords, const int* __restrict__ count, SpDims ssrc, const int* __restrict__ idx_src, int* __restrict__ tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = t / 27, o = t - r * 27;
  if (r >= *count) return;
  const int kz = o / 9, ky = (o / 3) % 3, kx = o % 3;
  const int z = coords[r * 3], y = coords[r * 3 + 1], x = coords[r * 3 + 2];
  int pz, py, px;
  bool ok = true;
  if (MODE == 0) { pz = z + kz - 1; py = y + ky - 1; px = x + kx - 1; }
  else if (MODE == 1) { pz = 2 * z - 1 + kz; py = 2 * y - 1 + ky; px = 2 * x - 1 + kx; }
  else {
    const int az = z + 1 - kz, ay = y + 1 - ky, ax = x + 1 - kx;
    ok = az >= 0 && ay >= 0 && ax >= 0 && !((az | ay | ax) & 1);
    pz = az >> 1; py = ay >> 1; px = ax >> 1;
  }
  tab[t] = (ok && sp_inside(ssrc, pz, py, px)) ? idx_src[sp_cell(ssrc, pz, py, px)] : -1;
}
// out = sum of the split-K partial tiles (split order: deterministic); scale_shift != NULL: evaluation-mode BatchNorm + ReLU on the sum
__global__ void k_sp_sum_parts(const float* __restrict__ part, int nsplit, size_t stride, const int* __restrict__ count, int cout,
                               const float* __restrict__ scale_shift, float* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(t / cout), c = (int)(t % cout);
  if (r >= *count) return;
  float v = part[t];
  for (int s = 1; s < nsplit; ++s) v += part[(size_t)s * stride + t];
  out[t] = scale_shift ? fmaxf(v * scale_shift[c] + scale_shift[cout + c], 0.f) : v;
}
constexpr int kSpSplitK = 3;

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

struct SpScratch { int* idx[4]; int* coords[4]; int* count; int* rowof; float* F[2]; float* Wt; float* ss;
                   int* tab_sub[4]; int* tab_down[3]; int* tab_downT[3]; float* part; float* canon; };
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
  for (int l = 0; l < 4; ++l) s.tab_sub[l] = (int*)take((size_t)p.cap[l] * 27 * sizeof(int));
  for (int l = 0; l < 3; ++l) {
    s.tab_down[l] = (int*)take((size_t)p.cap[l + 1] * 27 * sizeof(int));
    s.tab_downT[l] = (int*)take((size_t)p.cap[l] * 27 * sizeof(int));
  }
  s.part = (float*)take((size_t)kSpSplitK * maxcap * 96 * sizeof(float));
  s.canon = (float*)take(spconv_canon_floats() * sizeof(float));
  return off + 256;
}

// one convolution on the tensor cores: out[rows of `lo`][c_out] = sum_o W[:, o, :] . in[neighbour o]  (raw, or BatchNorm(eval) + ReLU when ss != NULL)
static int sp_conv_umma(const SherfSparseConv& L, const SpPlan& p, const SpScratch& s, int lo, const int* tab, const float* in, const float* ss,
                        float* out, cudaStream_t st) {
  CanonLayer cl;
  int rc = run_pack_spconv(L.weight, L.c_out, L.c_in, 0, s.canon, cl, st);
  if (rc) return rc;
  rc = launch_umma_spconv(cl, in, L.c_in, tab, s.count + lo, p.cap[lo], s.part, kSpSplitK, st);
  if (rc) return rc;
  k_sp_sum_parts<<<ceil_div((int64_t)p.cap[lo] * L.c_out, 256), 256, 0, st>>>(s.part, kSpSplitK, (size_t)p.cap[lo] * L.c_out, s.count + lo, L.c_out, ss, out);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}
// neighbour tables of a level (SubM) / of a strided convolution (forward table on the coarse rows, transposed table on the fine rows)
static int sp_table_sub(const SpPlan& p, const SpScratch& s, int l, cudaStream_t st) {
  k_sp_table<0><<<ceil_div((int64_t)p.cap[l] * 27, 256), 256, 0, st>>>(s.coords[l], s.count + l, p.dims[l], s.idx[l], s.tab_sub[l]);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}
static int sp_table_down(const SpPlan& p, const SpScratch& s, int l, bool transposed_too, cudaStream_t st) {
  k_sp_table<1><<<ceil_div((int64_t)p.cap[l + 1] * 27, 256), 256, 0, st>>>(s.coords[l + 1], s.count + l + 1, p.dims[l], s.idx[l], s.tab_down[l]);
  SHERF_LAUNCH_CHECK();
  if (transposed_too) {
    k_sp_table<2><<<ceil_div((int64_t)p.cap[l] * 27, 256), 256, 0, st>>>(s.coords[l], s.count + l, p.dims[l + 1], s.idx[l + 1], s.tab_downT[l]);
    SHERF_LAUNCH_CHECK();
  }
  return SHERF_OK;
}
static bool sp_simt() { const char* e = getenv("SHERF_SP_SIMT"); return e && e[0] == '1'; }      // the warp-per-row fp32 FMA kernels (first version; anchor)


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
  const bool simt = sp_simt();
  if (!simt) { const int rc = sp_table_sub(p, s, 0, st); if (rc) return rc; }
  // execution order and the levels emitted after conv1 / conv2 / conv3: renderer.py:756-782 (num_layers = 4)
  static const int emit_after[SHERF_SPARSE_CONVS] = {0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int c = 0; c < SHERF_SPARSE_CONVS; ++c) {
    const SherfSparseConv& L = enc.conv[c];
    if (L.c_in > 96 || L.c_out > 96 || L.c_in <= 0 || L.c_out <= 0 || (!simt && (L.c_in % 32 || L.c_out % 16))) {
      set_error("sparse conv %d: unsupported channels %d -> %d", c, L.c_in, L.c_out); return SHERF_E_UNSUPPORTED;
    }
    k_sp_pack<<<ceil_div(27 * L.c_in * L.c_out, 256), 256, 0, st>>>(L.weight, L.bn_weight, L.bn_bias, L.bn_mean, L.bn_var, L.c_in, L.c_out, s.Wt, s.ss);
    SHERF_LAUNCH_CHECK();
    if (L.kind == 0) {
      if (simt) {
        k_sp_conv<false><<<ceil_div(p.cap[level], 8), 256, 0, st>>>(s.coords[level], s.count + level, p.dims[level], s.idx[level], s.F[cur], L.c_in,
                                                                    L.c_out, s.Wt, s.ss, s.F[cur ^ 1]);
        SHERF_LAUNCH_CHECK();
      } else {
        const int rc = sp_conv_umma(L, p, s, level, s.tab_sub[level], s.F[cur], s.ss, s.F[cur ^ 1], st);
        if (rc) return rc;
      }
    } else {
      if (level >= 3) { set_error("sparse encoder: too many strided convs"); return SHERF_E_INVALID; }
      k_sp_down_sites<<<ceil_div((int64_t)p.cap[level] * 27, 256), 256, 0, st>>>(s.coords[level], s.count + level, p.dims[level + 1], s.idx[level + 1],
                                                                               s.count + level + 1, s.coords[level + 1], p.cap[level + 1]);
      SHERF_LAUNCH_CHECK();
      if (simt) {
        k_sp_conv<true><<<ceil_div(p.cap[level + 1], 8), 256, 0, st>>>(s.coords[level + 1], s.count + level + 1, p.dims[level], s.idx[level], s.F[cur],
                                                                       L.c_in, L.c_out, s.Wt, s.ss, s.F[cur ^ 1]);
        SHERF_LAUNCH_CHECK();
      } else {
        int rc = sp_table_down(p, s, level, false, st);
        if (!rc) rc = sp_conv_umma(L, p, s, level + 1, s.tab_down[level], s.F[cur], s.ss, s.F[cur ^ 1], st);
        if (!rc) rc = sp_table_sub(p, s, level + 1, st);
        if (rc) return rc;
      }
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


// =====================================================================================================================================
// Training mode: BatchNorm1d with BATCH statistics (nn.BatchNorm1d(eps=1e-3, momentum=0.01) in train(), renderer.py
// :822,840,...) and the backward pass torch.autograd derives through SparseConvNet.forward (renderer.py:744-785) when loss.backward() runs
// (loss.py:175).  Every activation is kept in the arena between the two calls.
//
// Rows and duplicates: spconv keeps one feature row per INPUT index, duplicates included, on the level-0 layers (SubMConv3d: output rows =
// input rows), and BatchNorm1d normalises over those rows.  A duplicate row carries the same value as its voxel's representative after the
// first convolution, so the statistics of the two level-0 BatchNorms weight every voxel by its multiplicity m (number of vertices in it);
// after down0 rows are unique output sites (m = 1).  The duplicate rows' outputs are never read downstream (hash tables and .dense() see the
// representative), but they do sit in the statistics, so in the backward pass a voxel hands  gamma rstd (dZ - m (dbeta + xhat dgamma) / N)
// to the convolution below it: the sum over its m rows, of which only the representative has an upstream gradient.
// =====================================================================================================================================

// level-0 multiplicity: vertices per voxel (after k_sp_index0: idx holds rows)
__global__ void k_sp_mult(const int* __restrict__ coord, int n, SpDims s, const int* __restrict__ idx, float* __restrict__ mult, int* __restrict__ nrows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = coord[i * 3], y = coord[i * 3 + 1], x = coord[i * 3 + 2];
  if (!sp_inside(s, z, y, x)) return;
  atomicAdd(&mult[idx[sp_cell(s, z, y, x)]], 1.f);
  atomicAdd(nrows, 1);
}

// per-channel sums over rows in double: part[split][c][2] = (sum m x, sum m x^2); lane = channel (coalesced rows), warp = row slice
constexpr int kSpStatSplits = 64;
__global__ void __launch_bounds__(256) k_sp_bn_stats(const float* __restrict__ raw, const int* __restrict__ count, int cout, const float* __restrict__ mult,
                                                     double* __restrict__ part) {
  __shared__ double red[8][32][2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int rows = *count;
  double s1 = 0.0, s2 = 0.0;
  if (c < cout)
    for (int r = blockIdx.y * 8 + w; r < rows; r += kSpStatSplits * 8) {
      const double x = raw[(size_t)r * cout + c], m = mult ? mult[r] : 1.f;
      s1 += m * x; s2 += m * x * x;
    }
  red[w][lane][0] = s1; red[w][lane][1] = s2;
  __syncthreads();
  if (w == 0 && c < cout) {
    for (int k = 1; k < 8; ++k) { s1 += red[k][lane][0]; s2 += red[k][lane][1]; }
    part[((size_t)blockIdx.y * cout + c) * 2] = s1;
    part[((size_t)blockIdx.y * cout + c) * 2 + 1] = s2;
  }
}
// stats[c] = mean, stats[96 + c] = biased variance, stats[192 + c] = rstd;  nrows: level-0 layers count duplicate rows (sum of m)
__global__ void k_sp_bn_finalize(const double* __restrict__ part, int cout, const int* __restrict__ nrows, float* __restrict__ stats, float* __restrict__ out_stats,
                                 int* __restrict__ out_rows) {
  const int c = threadIdx.x;
  if (c >= cout) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < kSpStatSplits; ++k) { s1 += part[((size_t)k * cout + c) * 2]; s2 += part[((size_t)k * cout + c) * 2 + 1]; }
  const double n = (double)max(*nrows, 1);
  const double mean = s1 / n;
  double var = s2 / n - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[c] = (float)mean; stats[96 + c] = (float)var; stats[192 + c] = (float)(1.0 / sqrt(var + 1e-3));
  if (out_stats) { out_stats[c] = (float)mean; out_stats[96 + c] = (float)var; }
  if (out_rows && c == 0) *out_rows = *nrows;
}
// evaluation-mode statistics for a differentiable forward: mean / var = the running statistics (constants of the graph)
__global__ void k_sp_bn_running(const float* __restrict__ mean, const float* __restrict__ var, int cout, float* __restrict__ stats) {
  const int c = threadIdx.x;
  if (c >= cout) return;
  stats[c] = mean[c]; stats[96 + c] = var[c]; stats[192 + c] = 1.f / sqrtf(var[c] + 1e-3f);
}
// act = relu((raw - mean) rstd gamma + beta)
__global__ void k_sp_bn_apply(const float* __restrict__ raw, const int* __restrict__ count, int cout, const float* __restrict__ stats, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* __restrict__ act) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(t / cout), c = (int)(t % cout);
  if (r >= *count) return;
  act[t] = fmaxf((raw[t] - stats[c]) * stats[192 + c] * gamma[c] + beta[c], 0.f);
}

// dA[r][c] (+)= g_vol[c][cell(r)]: the adjoint of .dense()
__global__ void k_sp_densify_bwd(const int* __restrict__ coords, const int* __restrict__ count, SpDims s, const float* __restrict__ gvol, int C, int accum,
                                 float* __restrict__ dA) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(t / C), c = (int)(t % C);
  if (r >= *count) return;
  const float g = gvol ? gvol[(size_t)c * s.d[0] * s.d[1] * s.d[2] + sp_cell(s, coords[r * 3], coords[r * 3 + 1], coords[r * 3 + 2])] : 0.f;
  dA[t] = accum ? dA[t] + g : g;
}
// part[split][c][2] = (sum dZ, sum dZ xhat), dZ = dA (act > 0), xhat = (raw - mean) rstd
__global__ void __launch_bounds__(256) k_sp_bn_bwd_stats(const float* __restrict__ dA, const float* __restrict__ act, const float* __restrict__ raw,
                                                         const int* __restrict__ count, int cout, const float* __restrict__ stats, double* __restrict__ part) {
  __shared__ double red[8][32][2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int rows = *count;
  double s1 = 0.0, s2 = 0.0;
  if (c < cout) {
    const float mean = stats[c], rstd = stats[192 + c];
    for (int r = blockIdx.y * 8 + w; r < rows; r += kSpStatSplits * 8) {
      const size_t i = (size_t)r * cout + c;
      const float dz = act[i] > 0.f ? dA[i] : 0.f;
      s1 += dz; s2 += (double)dz * (double)((raw[i] - mean) * rstd);
    }
  }
  red[w][lane][0] = s1; red[w][lane][1] = s2;
  __syncthreads();
  if (w == 0 && c < cout) {
    for (int k = 1; k < 8; ++k) { s1 += red[k][lane][0]; s2 += red[k][lane][1]; }
    part[((size_t)blockIdx.y * cout + c) * 2] = s1;
    part[((size_t)blockIdx.y * cout + c) * 2 + 1] = s2;
  }
}
// dbeta / dgamma out (+ kept in dsum[c], dsum[96 + c] for the apply kernel)
__global__ void k_sp_bn_bwd_finalize(const double* __restrict__ part, int cout, float* __restrict__ dsum, float* __restrict__ g_gamma, float* __restrict__ g_beta) {
  const int c = threadIdx.x;
  if (c >= cout) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < kSpStatSplits; ++k) { s1 += part[((size_t)k * cout + c) * 2]; s2 += part[((size_t)k * cout + c) * 2 + 1]; }
  dsum[c] = (float)s1; dsum[96 + c] = (float)s2;
  if (g_beta) g_beta[c] = (float)s1;
  if (g_gamma) g_gamma[c] = (float)s2;
}
// dRaw = gamma rstd (dZ - m (dbeta + xhat dgamma) / N)
__global__ void k_sp_bn_bwd_apply(const float* __restrict__ dA, const float* __restrict__ act, const float* __restrict__ raw, const int* __restrict__ count, int cout,
                                  const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ dsum, const float* __restrict__ mult,
                                  const int* __restrict__ nrows, float* __restrict__ dRaw) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(t / cout), c = (int)(t % cout);
  if (r >= *count) return;
  const float inv_n = nrows ? 1.f / (float)max(*nrows, 1) : 0.f;      // nrows == NULL: statistics were constants (evaluation mode)
  const float xhat = (raw[t] - stats[c]) * stats[192 + c];
  const float dz = act[t] > 0.f ? dA[t] : 0.f;
  const float m = mult ? mult[r] : 1.f;
  dRaw[t] = gamma[c] * stats[192 + c] * (dz - m * (dsum[c] + xhat * dsum[96 + c]) * inv_n);
}

// dX[j][ci] = sum over the outputs r that read input row j through offset o of  sum_co W[co][o][ci] dRaw[r][co].  Gather form from the INPUT
// side (no atomics): one warp per input row, lanes = input channels (W is KRSC: ci contiguous).
//   SubM: output at q = p - (k - 1);  strided: output at q with 2 q - 1 + k = p
template <bool DOWN>
__global__ void __launch_bounds__(256) k_sp_conv_bwd_x(const int* __restrict__ coords_in, const int* __restrict__ count_in, SpDims sout,
                                                       const int* __restrict__ idx_out, const float* __restrict__ dRaw, int cin, int cout,
                                                       const float* __restrict__ W, float* __restrict__ dX) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= *count_in) return;
  const int z = coords_in[j * 3], y = coords_in[j * 3 + 1], x = coords_in[j * 3 + 2];
  float acc[3] = {0.f, 0.f, 0.f};
  for (int o = 0; o < 27; ++o) {
    const int kz = o / 9, ky = (o / 3) % 3, kx = o % 3;
    int qz, qy, qx;
    if (DOWN) {
      const int az = z + 1 - kz, ay = y + 1 - ky, ax = x + 1 - kx;
      if (az < 0 || ay < 0 || ax < 0 || ((az | ay | ax) & 1)) continue;
      qz = az >> 1; qy = ay >> 1; qx = ax >> 1;
    } else {
      qz = z - (kz - 1); qy = y - (ky - 1); qx = x - (kx - 1);
    }
    if (!sp_inside(sout, qz, qy, qx)) continue;
    const int r = idx_out[sp_cell(sout, qz, qy, qx)];
    if (r < 0) continue;
    const float* g = dRaw + (size_t)r * cout;
    for (int co = 0; co < cout; ++co) {
      const float gv = g[co];
      const float* w = W + ((size_t)co * 27 + o) * cin;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int ci = lane + 32 * t;
        if (ci < cin) acc[t] = fmaf(gv, w[ci], acc[t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int ci = lane + 32 * t;
    if (ci < cin) dX[(size_t)j * cin + ci] = acc[t];
  }
}

// dW[co][o][ci] += sum_r dRaw[r][co] x[nbr(r, o)][ci].  grid (27 offsets, row splits); a block stages 8 (row, neighbour) pairs in shared
// memory, every thread owns up to 36 (co, ci) entries of the cout x cin tile; one atomicAdd per entry and block at the end.
constexpr int kSpDwSplits = 48;
template <bool DOWN>
__global__ void __launch_bounds__(256) k_sp_conv_bwd_w(const int* __restrict__ coords_out, const int* __restrict__ count_out, SpDims sin,
                                                       const int* __restrict__ idx_in, const float* __restrict__ dRaw, const float* __restrict__ X, int cin,
                                                       int cout, float* __restrict__ dW) {
  __shared__ float sg[8][96], sx[8][96];
  __shared__ int sj[8];
  const int o = blockIdx.x, kz = o / 9, ky = (o / 3) % 3, kx = o % 3;
  const int rows = *count_out;
  const int per = (rows + kSpDwSplits - 1) / kSpDwSplits;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  const int tci = threadIdx.x & 31, tco = threadIdx.x >> 5;      // ci = tci + 32 a (a < 3), co = tco + 8 b (b < 12)
  float acc[3][12];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 12; ++b) acc[a][b] = 0.f;
  for (int rb = r0; rb < r1; rb += 8) {
    __syncthreads();
    if (threadIdx.x < 8) {
      const int r = rb + threadIdx.x;
      int j = -1;
      if (r < r1) {
        const int z = coords_out[r * 3], y = coords_out[r * 3 + 1], x = coords_out[r * 3 + 2];
        const int pz = DOWN ? 2 * z - 1 + kz : z + kz - 1, py = DOWN ? 2 * y - 1 + ky : y + ky - 1, px = DOWN ? 2 * x - 1 + kx : x + kx - 1;
        if (sp_inside(sin, pz, py, px)) j = idx_in[sp_cell(sin, pz, py, px)];
      }
      sj[threadIdx.x] = j;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 96; i += 256) {
      const int q = i / 96, c = i - q * 96;
      const int j = sj[q];
      sg[q][c] = (j >= 0 && c < cout) ? dRaw[(size_t)(rb + q) * cout + c] : 0.f;
      sx[q][c] = (j >= 0 && c < cin) ? X[(size_t)j * cin + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (sj[q] < 0) continue;
      float xv[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) xv[a] = sx[q][tci + 32 * a];
#pragma unroll
      for (int b = 0; b < 12; ++b) {
        const float gv = sg[q][tco + 8 * b];
#pragma unroll
        for (int a = 0; a < 3; ++a) acc[a][b] = fmaf(gv, xv[a], acc[a][b]);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < 12; ++b) {
    const int co = tco + 8 * b;
    if (co >= cout) continue;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ci = tci + 32 * a;
      if (ci < cin && acc[a][b] != 0.f) atomicAdd(&dW[((size_t)co * 27 + o) * cin + ci], acc[a][b]);
    }
  }
}

// g_feat[i] = dX0[row of vertex i] for the representative vertex of a voxel, 0 for the others (their rows are never read, renderer.py:756)
__global__ void k_sp_feat_grad(const int* __restrict__ rowof, int n, int C, const float* __restrict__ dX, float* __restrict__ g_feat) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(t / C), c = (int)(t % C);
  if (i >= n) return;
  const int row = rowof[i];
  g_feat[t] = row >= 0 ? dX[(size_t)row * C + c] : 0.f;
}

// ---- arena of the training pass: the inference scratch + every layer's pre-BatchNorm output and activation + statistics ----
struct SpTrain {
  SpScratch s;
  float* raw[SHERF_SPARSE_CONVS]; float* act[SHERF_SPARSE_CONVS];
  float* stats;          // [13][288]: mean | biased var | rstd
  float* mult;           // [cap0]
  int* nrows;            // [13] rows under each BatchNorm (level 0: duplicates counted)
  int* lvl_of;           // host-side only (not carved)
  double* part;          // [kSpStatSplits][96][2]
  float* dsum;           // [192]
  float* dA; float* dB;  // [maxcap][96] gradient ping-pong
};
static const int kSpEmitAfter[SHERF_SPARSE_CONVS] = {0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
static const int kSpKind[SHERF_SPARSE_CONVS] = {0, 0, 1, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
// output level of conv c
static int sp_out_level(int c) { int l = 0; for (int i = 0; i <= c; ++i) l += kSpKind[i]; return l; }

static size_t sp_carve_train(char* base, int n, const SpPlan& p, const SherfSparseEncoder* enc, SpTrain& t) {
  size_t off = sp_carve(base, n, p, t.s);
  auto take = [&](size_t bytes) { off = (off + 255) & ~(size_t)255; char* q = base ? base + off : nullptr; off += bytes; return q; };
  int maxcap = 0;
  for (int l = 0; l < 4; ++l) if (p.cap[l] > maxcap) maxcap = p.cap[l];
  for (int c = 0; c < SHERF_SPARSE_CONVS; ++c) {
    const int cout = enc ? enc->conv[c].c_out : 96;
    const size_t rows = (size_t)p.cap[sp_out_level(c)];
    t.raw[c] = (float*)take(rows * cout * sizeof(float));
    t.act[c] = (float*)take(rows * cout * sizeof(float));
  }
  t.stats = (float*)take((size_t)SHERF_SPARSE_CONVS * 288 * sizeof(float));
  t.mult = (float*)take((size_t)p.cap[0] * sizeof(float));
  t.nrows = (int*)take(SHERF_SPARSE_CONVS * sizeof(int));
  t.part = (double*)take((size_t)kSpStatSplits * 96 * 2 * sizeof(double));
  t.dsum = (float*)take(192 * sizeof(float));
  t.dA = (float*)take((size_t)maxcap * 96 * sizeof(float));
  t.dB = (float*)take((size_t)maxcap * 96 * sizeof(float));
  return off + 256;
}

size_t sparse_encoder_train_scratch_bytes(int n, const int32_t* out_sh) {
  SpPlan p; sp_plan(n, out_sh, p);
  SpTrain t;
  return sp_carve_train(nullptr, n, p, nullptr, t) + 256;
}

static int sp_check_layers(const SherfSparseEncoder& enc) {
  for (int c = 0; c < SHERF_SPARSE_CONVS; ++c) {
    const SherfSparseConv& L = enc.conv[c];
    if (L.c_in > 96 || L.c_out > 96 || L.c_in <= 0 || L.c_out <= 0 || L.kind != kSpKind[c]) {
      set_error("sparse conv %d: unsupported layer (channels %d -> %d, kind %d)", c, L.c_in, L.c_out, L.kind);
      return SHERF_E_UNSUPPORTED;
    }
  }
  return SHERF_OK;
}

int run_sparse_encode_train(const SherfSparseEncoder& enc, const int* coord, const float* feat, int n, const int32_t* out_sh, float* const* vols,
                            float* batch_stats, int* row_counts, int use_running_stats, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  { const int rc = sp_check_layers(enc); if (rc) return rc; }
  SpPlan p; sp_plan(n, out_sh, p);
  char* base = (char*)scratch;
  const size_t mis = ((size_t)base) & 255;
  if (mis) base += 256 - mis;
  SpTrain t;
  const size_t need = sp_carve_train(base, n, p, &enc, t);
  if (need + 256 > scratch_bytes) { set_error("sparse-encoder training scratch too small: need %zu bytes, have %zu", need + 256, scratch_bytes); return SHERF_E_SCRATCH; }
  SpScratch& s = t.s;
  for (int l = 0; l < 4; ++l) SHERF_CUDA_OK(cudaMemsetAsync(s.idx[l], 0xff, p.cells[l] * sizeof(int), st));
  SHERF_CUDA_OK(cudaMemsetAsync(s.count, 0, 4 * sizeof(int), st));
  SHERF_CUDA_OK(cudaMemsetAsync(t.mult, 0, (size_t)p.cap[0] * sizeof(float), st));
  SHERF_CUDA_OK(cudaMemsetAsync(t.nrows, 0, SHERF_SPARSE_CONVS * sizeof(int), st));
  const int C0 = enc.conv[0].c_in;
  k_sp_claim<<<ceil_div(n, 256), 256, 0, st>>>(coord, n, p.dims[0], s.idx[0]);
  SHERF_LAUNCH_CHECK();
  k_sp_rows0<<<ceil_div(n, 256), 256, 0, st>>>(coord, feat, n, C0, p.dims[0], s.idx[0], s.count, s.rowof, s.coords[0], s.F[0]);
  SHERF_LAUNCH_CHECK();
  k_sp_index0<<<ceil_div(n, 256), 256, 0, st>>>(coord, n, p.dims[0], s.rowof, s.idx[0]);
  SHERF_LAUNCH_CHECK();
  k_sp_mult<<<ceil_div(n, 256), 256, 0, st>>>(coord, n, p.dims[0], s.idx[0], t.mult, t.nrows);      // nrows[0] = vertices inside the grid
  SHERF_LAUNCH_CHECK();
  int level = 0, emitted = 0;
  const float* in = s.F[0];
  // The forward that a backward pass follows keeps the fp32 FMA convolutions: it decides the ReLU gates, and the 3xTF32 tensor-core
  // products (2^-21 relative, 5e-6 on these outputs instead of 1e-6) flip ~ 5 x more near-zero units against the reference (measured,
  // tests/test_sparse_encoder.py).  The tensor cores run the evaluation forward (run_sparse_encode) and the input-gradient convolutions of
  // the backward, whose neighbour tables are built here.
  const bool tables = !sp_simt();
  if (tables) { const int rc = sp_table_sub(p, s, 0, st); if (rc) return rc; }
  for (int c = 0; c < SHERF_SPARSE_CONVS; ++c) {
    const SherfSparseConv& L = enc.conv[c];
    k_sp_pack<<<ceil_div(27 * L.c_in * L.c_out, 256), 256, 0, st>>>(L.weight, L.bn_weight, L.bn_bias, L.bn_mean, L.bn_var, L.c_in, L.c_out, s.Wt, s.ss);
    SHERF_LAUNCH_CHECK();
    if (L.kind == 0) {
      k_sp_conv<false><<<ceil_div(p.cap[level], 8), 256, 0, st>>>(s.coords[level], s.count + level, p.dims[level], s.idx[level], in, L.c_in, L.c_out,
                                                                  s.Wt, nullptr, t.raw[c]);
      SHERF_LAUNCH_CHECK();
    } else {
      k_sp_down_sites<<<ceil_div((int64_t)p.cap[level] * 27, 256), 256, 0, st>>>(s.coords[level], s.count + level, p.dims[level + 1], s.idx[level + 1],
                                                                               s.count + level + 1, s.coords[level + 1], p.cap[level + 1]);
      SHERF_LAUNCH_CHECK();
      k_sp_conv<true><<<ceil_div(p.cap[level + 1], 8), 256, 0, st>>>(s.coords[level + 1], s.count + level + 1, p.dims[level], s.idx[level], in, L.c_in,
                                                                     L.c_out, s.Wt, nullptr, t.raw[c]);
      SHERF_LAUNCH_CHECK();
      if (tables) {
        int rc = sp_table_down(p, s, level, true, st);
        if (!rc) rc = sp_table_sub(p, s, level + 1, st);
        if (rc) return rc;
      }
      ++level;
    }
    // rows under this BatchNorm: level 0 counts the duplicate rows (nrows[0], set by k_sp_mult), the other levels the unique output sites
    const int* nr = level == 0 ? t.nrows : s.count + level;
    if (use_running_stats) {
      k_sp_bn_running<<<1, 96, 0, st>>>(L.bn_mean, L.bn_var, L.c_out, t.stats + (size_t)c * 288);
      SHERF_LAUNCH_CHECK();
    } else {
      k_sp_bn_stats<<<dim3(ceil_div(L.c_out, 32), kSpStatSplits), 256, 0, st>>>(t.raw[c], s.count + level, L.c_out, level == 0 ? t.mult : nullptr, t.part);
      SHERF_LAUNCH_CHECK();
      k_sp_bn_finalize<<<1, 96, 0, st>>>(t.part, L.c_out, nr, t.stats + (size_t)c * 288, batch_stats ? batch_stats + (size_t)c * 192 : nullptr,
                                         row_counts ? row_counts + c : nullptr);
      SHERF_LAUNCH_CHECK();
    }
    k_sp_bn_apply<<<ceil_div((int64_t)p.cap[level] * L.c_out, 256), 256, 0, st>>>(t.raw[c], s.count + level, L.c_out, t.stats + (size_t)c * 288, L.bn_weight,
                                                                               L.bn_bias, t.act[c]);
    SHERF_LAUNCH_CHECK();
    in = t.act[c];
    if (kSpEmitAfter[c]) {
      SHERF_CUDA_OK(cudaMemsetAsync(vols[emitted], 0, p.cells[level] * (size_t)L.c_out * sizeof(float), st));
      k_sp_densify<<<ceil_div((int64_t)p.cap[level] * L.c_out, 256), 256, 0, st>>>(s.coords[level], s.count + level, p.dims[level], t.act[c], L.c_out,
                                                                                 vols[emitted]);
      SHERF_LAUNCH_CHECK();
      ++emitted;
    }
  }
  return SHERF_OK;
}

// SHERF_SP_DEBUG=1: per layer, sums of the incoming gradient, of its gated part and the number of open gates (host printout; diagnostics)
__global__ void k_sp_debug_sums(const float* __restrict__ dA, const float* __restrict__ act, const int* __restrict__ count, int cout, double* __restrict__ out) {
  double a = 0, b = 0, g = 0, n = 0;
  const int total = *count * cout;
  for (int i = threadIdx.x; i < total; i += blockDim.x) { a += dA[i]; b += fabs((double)dA[i]); if (act[i] > 0.f) { g += dA[i]; n += 1; } }
  __shared__ double red[4][256];
  red[0][threadIdx.x] = a; red[1][threadIdx.x] = b; red[2][threadIdx.x] = g; red[3][threadIdx.x] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) { double s = 0; for (int i = 0; i < 256; ++i) s += red[k][i]; out[k] = s; }
    out[4] = *count;
  }
}

// scratch: the arena run_sparse_encode_train filled (same n / out_sh / encoder), untouched in between
int run_sparse_encode_backward(const SherfSparseEncoder& enc, const int* coord, int n, const int32_t* out_sh, const float* const* g_vols,
                               const SherfSparseEncoderGrads& gr, float* g_feat, int use_running_stats, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  { const int rc = sp_check_layers(enc); if (rc) return rc; }
  SpPlan p; sp_plan(n, out_sh, p);
  char* base = (char*)scratch;
  const size_t mis = ((size_t)base) & 255;
  if (mis) base += 256 - mis;
  SpTrain t;
  const size_t need = sp_carve_train(base, n, p, &enc, t);
  if (need + 256 > scratch_bytes) { set_error("sparse-encoder training scratch too small: need %zu bytes, have %zu", need + 256, scratch_bytes); return SHERF_E_SCRATCH; }
  SpScratch& s = t.s;
  float* dA = t.dA;          // gradient w.r.t. the activation of layer c (rows of its output level)
  float* dB = t.dB;          // dRaw of layer c, then (in dA's place) the gradient w.r.t. its input
  bool have = false;         // dA holds the downstream convolution's input gradient
  for (int c = SHERF_SPARSE_CONVS - 1; c >= 0; --c) {
    const SherfSparseConv& L = enc.conv[c];
    const int lo = sp_out_level(c), li = lo - kSpKind[c];
    const int64_t elems = (int64_t)p.cap[lo] * L.c_out;
    if (kSpEmitAfter[c]) {
      const int e = lo - 1;                                           // levels 1, 2, 3 are emitted as volumes 0, 1, 2
      k_sp_densify_bwd<<<ceil_div(elems, 256), 256, 0, st>>>(s.coords[lo], s.count + lo, p.dims[lo], g_vols[e], L.c_out, have ? 1 : 0, dA);
      SHERF_LAUNCH_CHECK();
      have = true;
    }
    if (!have) { set_error("sparse encoder backward: no gradient reaches layer %d", c); return SHERF_E_INVALID; }
    if (getenv("SHERF_SP_DEBUG")) {
      double h[5];
      k_sp_debug_sums<<<1, 256, 0, st>>>(dA, t.act[c], s.count + lo, L.c_out, t.part);
      cudaMemcpyAsync(h, t.part, sizeof(h), cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      fprintf(stderr, "[sp bwd] layer %2d rows %4.0f sum dA % .6e  sum|dA| %.6e  gated sum % .6e  open gates %.0f\n", c, h[4], h[0], h[1], h[2], h[3]);
    }
    const float* stats = t.stats + (size_t)c * 288;
    const int* nr = use_running_stats ? nullptr : (lo == 0 ? t.nrows : s.count + lo);
    k_sp_bn_bwd_stats<<<dim3(ceil_div(L.c_out, 32), kSpStatSplits), 256, 0, st>>>(dA, t.act[c], t.raw[c], s.count + lo, L.c_out, stats, t.part);
    SHERF_LAUNCH_CHECK();
    k_sp_bn_bwd_finalize<<<1, 96, 0, st>>>(t.part, L.c_out, t.dsum, gr.bn_weight[c], gr.bn_bias[c]);
    SHERF_LAUNCH_CHECK();
    k_sp_bn_bwd_apply<<<ceil_div(elems, 256), 256, 0, st>>>(dA, t.act[c], t.raw[c], s.count + lo, L.c_out, stats, L.bn_weight, t.dsum,
                                                            lo == 0 ? t.mult : nullptr, nr, dB);
    SHERF_LAUNCH_CHECK();
    const float* X = c == 0 ? s.F[0] : t.act[c - 1];                  // the layer's input rows (level li)
    if (gr.weight[c]) {
      SHERF_CUDA_OK(cudaMemsetAsync(gr.weight[c], 0, (size_t)L.c_out * 27 * L.c_in * sizeof(float), st));
      if (L.kind == 0) k_sp_conv_bwd_w<false><<<dim3(27, kSpDwSplits), 256, 0, st>>>(s.coords[lo], s.count + lo, p.dims[li], s.idx[li], dB, X, L.c_in, L.c_out, gr.weight[c]);
      else k_sp_conv_bwd_w<true><<<dim3(27, kSpDwSplits), 256, 0, st>>>(s.coords[lo], s.count + lo, p.dims[li], s.idx[li], dB, X, L.c_in, L.c_out, gr.weight[c]);
      SHERF_LAUNCH_CHECK();
    }
    if (c > 0 || g_feat) {
      if (sp_simt()) {
        if (L.kind == 0) k_sp_conv_bwd_x<false><<<ceil_div(p.cap[li], 8), 256, 0, st>>>(s.coords[li], s.count + li, p.dims[lo], s.idx[lo], dB, L.c_in, L.c_out, L.weight, dA);
        else k_sp_conv_bwd_x<true><<<ceil_div(p.cap[li], 8), 256, 0, st>>>(s.coords[li], s.count + li, p.dims[lo], s.idx[lo], dB, L.c_in, L.c_out, L.weight, dA);
        SHERF_LAUNCH_CHECK();
      } else {
        // the adjoint is again a gathered linear layer: rows = the layer's INPUT sites, K = 27 c_out, weights with the two channel axes
        // swapped (and the offsets mirrored for SubMConv3d, whose forward table serves both directions)
        CanonLayer cl;
        int rc = run_pack_spconv(L.weight, L.c_out, L.c_in, L.kind == 0 ? 2 : 3, s.canon, cl, st);
        if (!rc) rc = launch_umma_spconv(cl, dB, L.c_out, L.kind == 0 ? s.tab_sub[li] : s.tab_downT[li], s.count + li, p.cap[li], s.part, kSpSplitK, st);
        if (rc) return rc;
        k_sp_sum_parts<<<ceil_div((int64_t)p.cap[li] * L.c_in, 256), 256, 0, st>>>(s.part, kSpSplitK, (size_t)p.cap[li] * L.c_in, s.count + li, L.c_in, nullptr, dA);
        SHERF_LAUNCH_CHECK();
      }
    }
  }
  if (g_feat) {
    k_sp_feat_grad<<<ceil_div((int64_t)n * enc.conv[0].c_in, 256), 256, 0, st>>>(s.rowof, n, enc.conv[0].c_in, dA, g_feat);
    SHERF_LAUNCH_CHECK();
  }
  return SHERF_OK;
}

}  // namespace sherf
