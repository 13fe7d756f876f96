"""TEST INFRASTRUCTURE -- not product code.  Only tests/, oracle/gen_golden.py and
bench.py's reference/cpu_baseline legs may import this.

Imports the reference's OWN `ImportanceRenderer` (renderer.py:260-704) and `NeRFDecoder`
(triplane.py:267-316), unmodified, from /root/reference under shims, so that golden vectors can be
generated from the real reference code on the CPU of THIS container.  It cannot
travel: /root/reference does not exist on the GPU box, where `available()` is False and the
restatement in oracle/port.py (validated against this module here) is the checker.

Shims (none of them touches reference arithmetic except #1, which replaces an absent dependency):
 1. pytorch3d.ops.knn.knn_points -> brute-force K=1 squared-L2 argmin, d2 = (dx*dx + dy*dy) + dz*dz
    in fp32, smallest index on ties.  pytorch3d is not vendored and its version is unpinned
    (README.md:48), so this formula IS our statement of its semantics ("parity unpinned").
 2. spconv / spconv.pytorch -> oracle/spconv_shim.py: stand-ins that hold a zero `weight` of spconv 2.3.3's shape and evaluate
    the three spconv ops the reference uses (rules restated there); the render fixtures still take the three densified volumes
    as inputs (shim 6), the reference's SparseConvNet.forward runs under them in tests/test_sparse_encoder.py.
 3. torch.Tensor.cuda -> identity, torch.cuda.current_device -> 0 (the reference hard-codes .cuda()).
 4. renderer.read_pickle / SMPL_to_tensor -> return the synthetic SMPL-shaped model.
 5. imageio -> empty stub (imported, unused, by triplane.py:27).
 6. ImportanceRenderer.encoder_3d -> a module doing exactly renderer.py:764,773,782,794-795 on the
    three supplied dense volumes.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

# the mounted reference where it exists (build container); else the byte-identical copies oracle/make_ref.py left in oracle/_ref/
# (git-ignored, shipped to the GPU box with the snapshot)
_MOUNTED = '/root/reference/sherf'
_COPIED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'sherf')
REF_ROOT = _MOUNTED if os.path.isdir(os.path.join(_MOUNTED, 'training', 'volumetric_rendering')) else _COPIED


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'training', 'volumetric_rendering'))


def mounted() -> bool:
    """True where the full reference tree is mounted (tests that need more than the hot-path closure skip otherwise)."""
    return REF_ROOT == _MOUNTED


def knn_points_bruteforce(p1, p2, K=1, chunk=None, **_):
    """Stand-in for pytorch3d.ops.knn.knn_points (call sites renderer.py:315,564,627)."""
    assert K == 1 and p1.shape[0] == 1 and p2.shape[0] == 1
    q, v = p1[0], p2[0]
    chunk = chunk or (65536 if q.is_cuda else 4096)          # [chunk, V] temporaries: 1.8 GB each on the GPU, 113 MB on the CPU
    vx, vy, vz = v[:, 0][None], v[:, 1][None], v[:, 2][None]
    d2_out = torch.empty(q.shape[0], dtype=torch.float32, device=q.device)
    id_out = torch.empty(q.shape[0], dtype=torch.long, device=q.device)
    for s in range(0, q.shape[0], chunk):
        c = q[s:s + chunk]
        dx = c[:, 0:1] - vx
        dy = c[:, 1:2] - vy
        dz = c[:, 2:3] - vz
        d2 = (dx * dx + dy * dy) + dz * dz
        m, i = torch.min(d2, dim=1)
        # torch.min returns *a* minimal index; force the smallest one for determinism
        first = (d2 == m[:, None]).float().argmax(dim=1)
        d2_out[s:s + chunk] = m
        id_out[s:s + chunk] = first
    return d2_out.view(1, -1, 1), id_out.view(1, -1, 1), None


class DenseVolumeGather(nn.Module):
    """Shim 6: the gather half of SparseConvNet.forward (renderer.py:762-797)."""

    def forward(self, volumes, grid):
        feats = [F.grid_sample(v, grid, padding_mode='zeros', align_corners=True) for v in volumes]
        feats = torch.cat(feats, dim=1)
        return feats.view(feats.size(0), -1, feats.size(4)).transpose(1, 2)


# Shim 2 (functional since round 2, oracle/spconv_shim.py): SubMConv3d / SparseConv3d / SparseSequential / SparseConvTensor stand-ins that
# hold `weight` in spconv 2.3.3's [out, k, k, k, in] layout AND evaluate the convolutions (dense formulation), so that the reference's own
# SparseConvNet.forward can run.  The old constructor-only names stay as aliases (tests/helpers/make_reference_snapshot.py subclasses them).
from oracle import spconv_shim                                            # noqa: E402
SpconvConvStub = spconv_shim.SparseConvolution
SpconvSequentialStub = spconv_shim.SparseSequential


_loaded = None


def load(smpl_model_torch: dict):
    """Returns (renderer_module, triplane_NeRFDecoder_class). Idempotent."""
    global _loaded
    if _loaded is not None:
        _loaded[0]._SYNTH_SMPL = smpl_model_torch
        return _loaded[0], _loaded[1]
    if not available():
        raise RuntimeError('reference tree not present (expected on the GPU box); use oracle.port')
    sys.dont_write_bytecode = True          # /root/reference is read-only

    knn_mod = types.ModuleType('pytorch3d.ops.knn'); knn_mod.knn_points = knn_points_bruteforce
    ops_mod = types.ModuleType('pytorch3d.ops'); ops_mod.knn = knn_mod
    p3d = types.ModuleType('pytorch3d'); p3d.ops = ops_mod
    sys.modules.update({'pytorch3d': p3d, 'pytorch3d.ops': ops_mod, 'pytorch3d.ops.knn': knn_mod})

    sp = types.ModuleType('spconv.pytorch')
    sp.SparseSequential, sp.SubMConv3d, sp.SparseConv3d = spconv_shim.SparseSequential, spconv_shim.SubMConv3d, spconv_shim.SparseConv3d
    sp.SparseConvTensor = spconv_shim.SparseConvTensor
    core = types.ModuleType('spconv.core'); core.SparseConvTensor = spconv_shim.SparseConvTensor
    sp.core = core
    spr = types.ModuleType('spconv'); spr.pytorch = sp; spr.core = core
    sys.modules.update({'spconv': spr, 'spconv.pytorch': sp, 'spconv.core': core})
    sys.modules.setdefault('imageio', types.ModuleType('imageio'))

    if not torch.cuda.is_available() or os.environ.get('SHERF_REF_FORCE_CPU') == '1':
        # shim 3: the reference hard-codes .cuda(); on a CPU-only box, or when its CPU path is being timed on a GPU box
        # (SHERF_REF_FORCE_CPU=1: bench.py --impl reference, always in its own process), .cuda() is the identity
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.current_device = lambda: 0

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from training.volumetric_rendering import renderer as ref_renderer   # noqa: E402
    ref_renderer._SYNTH_SMPL = smpl_model_torch
    ref_renderer.read_pickle = lambda path: ref_renderer._SYNTH_SMPL
    ref_renderer.SMPL_to_tensor = lambda params, device: params

    # NeRFDecoder lives in triplane.py, whose module-level imports pull in the StyleGAN backbone and
    # torchvision; exec only the class body's source region is not allowed (no copying), so import it.
    try:
        from training import triplane as ref_triplane
        decoder_cls = ref_triplane.NeRFDecoder
    except Exception as e:                                               # pragma: no cover
        raise RuntimeError(f'could not import reference triplane.py under shims: {e}')
    _loaded = (ref_renderer, decoder_cls)
    return _loaded


def build_reference(smpl_model_torch: dict, seed: int = 0):
    """Reference ImportanceRenderer(use_trans=True, use_NeRF_decoder=True) + NeRFDecoder(32), default torch init under `seed`."""
    ref_renderer, decoder_cls = load(smpl_model_torch)
    torch.manual_seed(seed)
    ren = ref_renderer.ImportanceRenderer(use_1d_feature=True, use_2d_feature=True, use_3d_feature=True,
                                          use_trans=True, use_NeRF_decoder=True)
    ren.encoder_3d = DenseVolumeGather()
    dec = decoder_cls(32)
    return ren.eval(), dec.eval()


@torch.no_grad()
def render(ren, dec, scene: dict):
    """One call of the reference forward on a synthetic scene (synthetic.make_scene)."""
    return ren(scene['planes'], scene['obs_input_img'], scene['obs_input_feature'], scene['volumes'], None,
               scene['obs_sp_input'], dec, scene['ray_origins'], scene['ray_directions'], scene['near'], scene['far'],
               scene['input_data'], scene['rendering_options'])


@torch.no_grad()
def render_importance(ren, dec, scene: dict, u: torch.Tensor, return_stages: bool = False):
    """The REPAIRED coarse+fine forward, composed only of the reference's own methods.

    The reference's fine pass (renderer.py:375-393) cannot run: :376 calls the ray marcher without `rays_d` and :383
    calls `run_model` with the 5-argument EG3D signature.  Repair = fix those two call sites and nothing else:
      * coarse colours / densities / depths: the reference's forward (:299-371), captured at its ray-marcher call;
      * coarse weights: the reference's MipRayMarcher2 with `ray_directions` passed (:376 repaired);
      * fine depths: the reference's sample_importance / sample_pdf (:483-542), with torch.rand (:526) returning `u`;
      * fine colours / densities: the reference's forward again with sample_stratified returning the fine depths
        (:383 repaired: the fine points go through the same cull, warps, gathers and decoder, :304-371);
      * the reference's unify_samples (:446-456) and ray marcher (:393).
    """
    opts = scene['rendering_options']
    n_imp = int(opts['depth_resolution_importance'])
    opts0 = dict(opts, depth_resolution_importance=0)
    marcher = ren.ray_marcher
    cap = []

    class _Tap(nn.Module):
        def forward(self, colors, densities, depths, rays_d, options):
            cap.append((colors, densities, depths))
            return marcher(colors, densities, depths, rays_d, options)

    def fwd():
        return ren(scene['planes'], scene['obs_input_img'], scene['obs_input_feature'], scene['volumes'], None,
                   scene['obs_sp_input'], dec, scene['ray_origins'], scene['ray_directions'], scene['near'], scene['far'],
                   scene['input_data'], opts0)
    ren.ray_marcher = _Tap()
    orig_rand = torch.rand
    try:
        fwd()
        colors_c, dens_c, depths_c = cap[0]
        weights_c = marcher(colors_c, dens_c, depths_c, scene['ray_directions'], opts)[2]
        torch.rand = lambda *a, **k: u
        try:
            depths_f = ren.sample_importance(depths_c, weights_c, n_imp)
        finally:
            torch.rand = orig_rand
        ren.sample_stratified = lambda *a, **k: depths_f
        try:
            fwd()
        finally:
            del ren.sample_stratified
        colors_f, dens_f, _ = cap[1]
    finally:
        ren.ray_marcher = marcher
    all_d, all_c, all_s = ren.unify_samples(depths_c, colors_c, dens_c, depths_f, colors_f, dens_f)
    rgb, depth, w = marcher(all_c, all_s, all_d, scene['ray_directions'], opts)
    out = (rgb, depth, w.sum(2))
    if return_stages:
        return out + ({'coarse_weights': weights_c[0, :, :, 0], 't_fine': depths_f[0, :, :, 0], 'sigma_fine': dens_f[0, :, :, 0],
                       'colors_fine': colors_f[0], 'sigma_coarse': dens_c[0, :, :, 0]},)
    return out
