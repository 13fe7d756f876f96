// Dataset-side SMPL forward on the device: posed vertices of one frame, what the datasets compute on the host
// with sherf/smpl/smpl_numpy.py:46-98 (`SMPL.__call__`) followed by `xyz @ R.T + Th` (RenderPeople_dataset.py:210).  With
// sherf_generate_rays this removes the last per-frame host product of a streamed sequence: only pose / shape / camera are uploaded.
// Arithmetic in fp64 like numpy (the model arrays are float64 there; the rotation matrices are ROUNDED TO float32 exactly where
// cv2.Rodrigues returns float32, smpl_numpy.py:61-65), float32 out.
#include "common.cuh"
#include "stages.cuh"

namespace sherf {

struct SmplFwdScratch {
  double J[kJoints * 3];        // rest joints of the shaped template
  double G[kJoints * 12];       // 3x4 rows of the skinning transforms (rest-joint correction applied)
  double lrot[kPoseFeat];       // (R[1:] - I) flattened
  double joints[kJoints * 3];   // posed joint positions G[:, :3, 3] before the correction
};

// block j: J[j] = J_regressor[j,:] . (v_template + shapedirs . beta)
__global__ void __launch_bounds__(256) k_smplf_joints(const float* __restrict__ v_template, const float* __restrict__ shapedirs,
                                                      const float* __restrict__ j_regressor, const float* __restrict__ beta, int V,
                                                      SmplFwdScratch* __restrict__ sc) {
  __shared__ double red[3][256];
  __shared__ double sb[10];
  const int j = blockIdx.x;
  if (threadIdx.x < 10) sb[threadIdx.x] = (double)beta[threadIdx.x];
  __syncthreads();
  double a[3] = {0.0, 0.0, 0.0};
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const double w = (double)j_regressor[(size_t)j * V + v];
    if (w != 0.0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 10; ++k) s += (double)shapedirs[((size_t)v * 3 + c) * 10 + k] * sb[k];
        a[c] += w * (s + (double)v_template[v * 3 + c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) red[c][threadIdx.x] = a[c];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
#pragma unroll
      for (int c = 0; c < 3; ++c) red[c][threadIdx.x] += red[c][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 3) sc->J[j * 3 + threadIdx.x] = red[threadIdx.x][0];
}

struct ParentsD { int p[kJoints]; };

// one thread: Rodrigues (cv2 semantics: double arithmetic, float32 result), kinematic chain, rest-joint correction
__global__ void k_smplf_chain(const float* __restrict__ poses, ParentsD par, SmplFwdScratch* __restrict__ sc) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double R[kJoints][9];
  for (int j = 0; j < kJoints; ++j) {
    const double rx = (double)poses[j * 3], ry = (double)poses[j * 3 + 1], rz = (double)poses[j * 3 + 2];
    const double theta = sqrt(rx * rx + ry * ry + rz * rz);
    double M[9];
    if (theta < 2.220446049250313e-16) {
      for (int e = 0; e < 9; ++e) M[e] = (e % 4 == 0) ? 1.0 : 0.0;
    } else {
      const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
      const double x = rx * it, y = ry * it, z = rz * it;
      M[0] = c + c1 * x * x;     M[1] = c1 * x * y - s * z; M[2] = c1 * x * z + s * y;
      M[3] = c1 * x * y + s * z; M[4] = c + c1 * y * y;     M[5] = c1 * y * z - s * x;
      M[6] = c1 * x * z - s * y; M[7] = c1 * y * z + s * x; M[8] = c + c1 * z * z;
    }
    for (int e = 0; e < 9; ++e) R[j][e] = (double)(float)M[e];            // cv2.Rodrigues hands back float32 for a float32 vector
  }
  for (int j = 1; j < kJoints; ++j)
    for (int e = 0; e < 9; ++e) sc->lrot[(j - 1) * 9 + e] = (double)((float)R[j][e] - ((e % 4 == 0) ? 1.0f : 0.0f));   // float32 subtraction (smpl_numpy.py:70-71)
  double G[kJoints][12];
  for (int j = 0; j < kJoints; ++j) {
    double L[12];
    const int pj = j == 0 ? 0 : par.p[j];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) L[r * 4 + c] = R[j][r * 3 + c];
      L[r * 4 + 3] = sc->J[j * 3 + r] - (j == 0 ? 0.0 : sc->J[pj * 3 + r]);
    }
    if (j == 0) { for (int e = 0; e < 12; ++e) G[0][e] = L[e]; }
    else
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
          double s = 0.0;
          for (int k = 0; k < 3; ++k) s += G[pj][r * 4 + k] * L[k * 4 + c];
          G[j][r * 4 + c] = s + (c == 3 ? G[pj][r * 4 + 3] : 0.0);
        }
  }
  for (int j = 0; j < kJoints; ++j) {
    for (int r = 0; r < 3; ++r) {
      sc->joints[j * 3 + r] = G[j][r * 4 + 3];
      double corr = 0.0;
      for (int k = 0; k < 3; ++k) corr += G[j][r * 4 + k] * sc->J[j * 3 + k];
      for (int c = 0; c < 3; ++c) sc->G[j * 12 + r * 4 + c] = G[j][r * 4 + c];
      sc->G[j * 12 + r * 4 + 3] = G[j][r * 4 + 3] - corr;
    }
  }
}

// thread per vertex: v_posed = v_shaped + posedirs . lrot; T = sum_j w_j G_j; v = T [v_posed, 1]; optional world transform v R^T + Th
__global__ void __launch_bounds__(128) k_smplf_skin(const float* __restrict__ v_template, const float* __restrict__ shapedirs,
                                                    const float* __restrict__ posedirs, const float* __restrict__ weights,
                                                    const float* __restrict__ beta, const SmplFwdScratch* __restrict__ sc, int V,
                                                    const float* __restrict__ Rw, const float* __restrict__ Thw, float* __restrict__ verts_smpl,
                                                    float* __restrict__ verts_world) {
  __shared__ double sG[kJoints * 12], sl[kPoseFeat], sb[10];
  for (int i = threadIdx.x; i < kJoints * 12; i += blockDim.x) sG[i] = sc->G[i];
  for (int i = threadIdx.x; i < kPoseFeat; i += blockDim.x) sl[i] = sc->lrot[i];
  if (threadIdx.x < 10) sb[threadIdx.x] = (double)beta[threadIdx.x];
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  double vp[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 10; ++k) s += (double)shapedirs[((size_t)v * 3 + c) * 10 + k] * sb[k];
    double o = 0.0;
    const float* pr = posedirs + ((size_t)v * 3 + c) * kPoseFeat;
    for (int k = 0; k < kPoseFeat; ++k) o += (double)pr[k] * sl[k];
    vp[c] = (s + (double)v_template[v * 3 + c]) + o;
  }
  double T[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) T[e] = 0.0;
  for (int j = 0; j < kJoints; ++j) {
    const double w = (double)weights[(size_t)v * kJoints + j];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] += w * sG[j * 12 + e];
  }
  double out[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) out[r] = ((T[r * 4] * vp[0] + T[r * 4 + 1] * vp[1]) + T[r * 4 + 2] * vp[2]) + T[r * 4 + 3];
  if (verts_smpl) { verts_smpl[v * 3] = (float)out[0]; verts_smpl[v * 3 + 1] = (float)out[1]; verts_smpl[v * 3 + 2] = (float)out[2]; }
  if (verts_world) {
    // xyz = (np.matmul(xyz, R.transpose()) + Th).astype(float32): xyz is float64 there, R / Th float32   (RenderPeople_dataset.py:210)
#pragma unroll
    for (int r = 0; r < 3; ++r)
      verts_world[v * 3 + r] = (float)(((out[0] * (double)Rw[r * 3] + out[1] * (double)Rw[r * 3 + 1]) + out[2] * (double)Rw[r * 3 + 2]) + (double)Thw[r]);
  }
}

size_t smpl_forward_scratch_bytes() { return sizeof(SmplFwdScratch) + 512; }

int run_smpl_vertices(const SherfSmplModel& smpl, const SherfPose& pose, float* verts_smpl, float* verts_world, void* scratch, size_t scratch_bytes,
                      cudaStream_t st) {
  if (scratch_bytes < smpl_forward_scratch_bytes()) { set_error("scratch arena too small for sherf_smpl_vertices"); return SHERF_E_SCRATCH; }
  char* b = (char*)scratch;
  const size_t mis = ((size_t)b) & 255;
  if (mis) b += 256 - mis;
  SmplFwdScratch* sc = (SmplFwdScratch*)b;
  const int V = smpl.n_verts;
  ParentsD par;
  for (int j = 0; j < kJoints; ++j) par.p[j] = j == 0 ? 0 : smpl.parents[j];
  k_smplf_joints<<<kJoints, 256, 0, st>>>(smpl.v_template, smpl.shapedirs, smpl.j_regressor, pose.shapes, V, sc);
  SHERF_LAUNCH_CHECK();
  k_smplf_chain<<<1, 32, 0, st>>>(pose.poses, par, sc);
  SHERF_LAUNCH_CHECK();
  k_smplf_skin<<<ceil_div(V, 128), 128, 0, st>>>(smpl.v_template, smpl.shapedirs, smpl.posedirs, smpl.weights, pose.shapes, sc, V, pose.R, pose.Th,
                                                verts_smpl, (verts_world && pose.R && pose.Th) ? verts_world : nullptr);
  SHERF_LAUNCH_CHECK();
  return SHERF_OK;
}

}  // namespace sherf
