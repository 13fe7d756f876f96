/* A plain-C client of include/sherf_b200.h: proves the header is valid C (no torch / C++ types at the boundary), that every entry
 * point a reference-side binding would use resolves from a dlopen'ed libsherf_b200.so, and that argument validation returns error
 * codes instead of crashing -- all without a GPU.  Built and run by tests/test_abi.py. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "sherf_b200.h"

#define LOAD(name) do { *(void**)(&p_##name) = dlsym(h, #name); if (!p_##name) { printf("missing %s\n", #name); return 2; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { printf("dlopen failed: %s\n", dlerror()); return 2; }
  int (*p_sherf_abi_version)(void);
  const char* (*p_sherf_last_error)(void);
  size_t (*p_sherf_scratch_bytes)(const SherfScene*, int32_t, int32_t, int32_t, int32_t);
  int (*p_sherf_render_forward)(const SherfSmplModel*, const SherfFrame*, const SherfScene*, const SherfWeights*, const SherfRays*,
                                const SherfOptions*, const SherfOut*, const SherfDebug*, void*, size_t, void*, int64_t*);
  int (*p_sherf_sparse_encode)(const SherfSparseEncoder*, const int32_t*, const float*, int32_t, const int32_t*, float*, float*, float*, void*,
                               size_t, void*);
  int (*p_sherf_generate_rays)(const double*, const double*, const double*, int32_t, int32_t, const double*, float*, float*, float*, float*,
                               uint8_t*, void*);
  LOAD(sherf_abi_version); LOAD(sherf_last_error); LOAD(sherf_scratch_bytes); LOAD(sherf_render_forward); LOAD(sherf_sparse_encode);
  LOAD(sherf_generate_rays);
  if (p_sherf_abi_version() != SHERF_ABI_VERSION) { printf("abi mismatch\n"); return 3; }
  SherfScene sc;
  memset(&sc, 0, sizeof sc);
  sc.plane_ch = 32; sc.plane_h = sc.plane_w = 256;
  sc.img_h = sc.img_w = 512;
  sc.feat_ch = 64; sc.feat_h = sc.feat_w = 256;
  {
    const int ch[3] = {32, 64, 96}, d[3][3] = {{48, 160, 192}, {24, 80, 96}, {12, 40, 48}};
    for (int l = 0; l < 3; ++l) { sc.vol_ch[l] = ch[l]; for (int a = 0; a < 3; ++a) sc.vol_dim[l][a] = d[l][a]; }
  }
  const size_t coarse = p_sherf_scratch_bytes(&sc, 512 * 512, 64, 0, 6890);
  const size_t both = p_sherf_scratch_bytes(&sc, 512 * 512, 64, 64, 6890);
  if (!(coarse > 0 && both > coarse)) { printf("scratch sizes %zu %zu\n", coarse, both); return 4; }
  if (p_sherf_render_forward(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) != SHERF_E_INVALID || !strstr(p_sherf_last_error(), "null")) return 5;
  if (p_sherf_sparse_encode(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) != SHERF_E_INVALID) return 6;
  if (p_sherf_generate_rays(0, 0, 0, 4, 4, 0, 0, 0, 0, 0, 0, 0) != SHERF_E_INVALID) return 7;
  printf("ok abi=%d scratch_coarse_MB=%zu scratch_coarse_fine_MB=%zu sizeof(SherfWeights)=%zu\n", p_sherf_abi_version(), coarse >> 20, both >> 20,
         sizeof(SherfWeights));
  dlclose(h);
  return 0;
}
