"""TEST INFRASTRUCTURE.  Golden vector of the generator-level boundary: runs the REFERENCE's own `TriPlaneGenerator.synthesis`
(triplane.py:81-172, unmodified, imported under oracle/ref_shim.py) on a seeded synthetic scene and writes
tests/golden/synthesis_<case>.npz -- the returned image dict plus the observation-side intermediates of triplane.py:105-137
(vertex features, visibility mask, canonical vertices, voxel coordinates, box).

    python -m oracle.gen_golden_synthesis

What stands in for code outside the hot-path scope (none of it is arithmetic of the path under test):
  * `self` of synthesis is a namespace carrying the reference's ImportanceRenderer / NeRFDecoder (ref_shim.build_reference), a seeded
    Conv1d(96,32,1), and two stand-ins that RETURN the scene's random tri-planes / 2-D feature map where the StyleGAN2 backbone and the
    ResNet-18 encoder would produce them (out of scope);
  * spconv.core.SparseConvTensor is a record; the renderer's `encoder_3d` evaluates oracle/sparse_encoder.py (spconv semantics stated
    there: "parity unpinned" for that sub-step) with seeded weights and then performs the reference's own dense gathers.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from sherf_b200 import synthetic as S
from oracle import port, ref_shim, sparse_encoder as SE

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
CASES = {
    'synthesis_32x32x16': dict(spec=S.SceneSpec(H=32, W=32, samples=16, seed=17), weight_seed=5, enc_seed=6, proj_seed=7),
}


class SparseRecord:
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features, self.indices, self.spatial_shape, self.batch_size = features, indices, spatial_shape, batch_size


def encoder_state(enc_seed):
    """Seeded weights of renderer.encoder_3d with non-trivial BatchNorm statistics; regenerated identically by the GPU test."""
    from sherf_b200.renderer import SparseConvNet
    torch.manual_seed(enc_seed)
    return SE.random_state_dict(SparseConvNet(4), enc_seed)


def state_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def projection_conv(seed):
    torch.manual_seed(seed)
    return nn.Conv1d(96, 32, 1)


def run_case(name, cfg, model, model_t):
    ren, dec = ref_shim.build_reference(model_t, cfg['weight_seed'])
    with torch.no_grad():
        dec.alpha_linear.weight *= 30
        dec.alpha_linear.bias += 2.0
    ref_tp = sys.modules['training.triplane']
    sys.modules['spconv.pytorch'].core.SparseConvTensor = SparseRecord
    scene = S.make_scene(cfg['spec'], model)
    enc_sd = encoder_state(cfg['enc_seed'])
    cap = {}

    class EncoderShim(nn.Module):
        def forward(self, sp, grid):
            cap['vert_feat'], cap['coord'], cap['out_sh'] = sp.features.clone(), sp.indices.clone(), list(sp.spatial_shape)
            vols = SE.encode_sparse(enc_sd, sp.indices[:, 1:], sp.features, sp.spatial_shape)
            return ref_shim.DenseVolumeGather()(vols, grid)
    ren.encoder_3d = EncoderShim()
    proj = projection_conv(cfg['proj_seed'])
    planes96 = scene['planes'].view(1, 96, 256, 256)
    fake = types.SimpleNamespace(
        renderer=ren, decoder=dec, conv1d_projection=proj, use_3d_feature=True, neural_rendering_resolution=64, _last_planes=None,
        rendering_kwargs=dict(scene['rendering_options']), superresolution=None,
        encoder_2d_feature=lambda img, extract_feature=False: scene['obs_input_feature'],
        backbone=types.SimpleNamespace(synthesis=lambda ws, update_emas=False, **k: planes96))
    orig_prep = ref_tp.TriPlaneGenerator.prepare_sp_input

    def prep(self, vertex, xyz):
        cap['can'] = xyz[0].clone()
        r = orig_prep(self, vertex, xyz)
        cap['bounds'] = r[0]['bounds'].clone()
        return r
    fake.prepare_sp_input = types.MethodType(prep, fake)
    o_proj = ren.projection

    def proj_tap(*a, **k):
        r = o_proj(*a, **k)
        if isinstance(r, tuple):
            cap['vmask'] = r[1].clone()
        return r
    ren.projection = proj_tap
    # compute_normal's `norm[:, faces[:, k]] += n` (renderer.py:58-60) is an index_put WITHOUT accumulation over repeated vertex ids:
    # which face wins is implementation-defined in torch (it differs between 1 and 8 CPU threads, and is undefined on CUDA).  The
    # fixture pins the sequential semantic -- the last face in index order wins, which is also numpy's -- by running single-threaded.
    nthreads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        with torch.no_grad():
            out = ref_tp.TriPlaneGenerator.synthesis(fake, None, scene['input_data'], None, use_sr_module=False, test_flag=True)
    finally:
        torch.set_num_threads(nthreads)
    ren.projection = o_proj
    assert cap['out_sh'] == scene['obs_sp_input']['out_sh'], (cap['out_sh'], scene['obs_sp_input']['out_sh'])
    spec = cfg['spec']
    res = {
        'spec': np.array([spec.H, spec.W, spec.samples, spec.seed, int(spec.random_global_R), int(spec.white_back)], np.int64),
        'weight_seed': np.int64(cfg['weight_seed']), 'enc_seed': np.int64(cfg['enc_seed']), 'proj_seed': np.int64(cfg['proj_seed']),
        'enc_sha256': np.array(state_checksum(enc_sd)),
        'image': out['image'].numpy(), 'image_raw': out['image_raw'].numpy(), 'image_depth': out['image_depth'].numpy(),
        'weights_image': out['weights_image'].numpy(),
        'vert_feat': cap['vert_feat'].numpy(), 'coord': cap['coord'].numpy().astype(np.int32), 'out_sh': np.array(cap['out_sh'], np.int32),
        'bounds': cap['bounds'].numpy(), 'can': cap['can'].numpy(), 'vmask': np.packbits(cap['vmask'].reshape(-1).numpy()),
        'proj_w': proj.weight.detach().numpy(), 'proj_b': proj.bias.detach().numpy(),
    }
    for name_w, t in port.hot_path_state_dict(ren, dec).items():
        res['w/' + name_w] = t.numpy()
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **res)
    dup = cap['coord'].shape[0] - len({tuple(c) for c in cap['coord'].tolist()})
    print(f'{name}: image {tuple(out["image"].shape)} acc.max={float(out["weights_image"].max()):.3f} visible vertices '
          f'{int(cap["vmask"].sum())}/{cap["vmask"].numel()} duplicate voxels {dup} out_sh={cap["out_sh"]} -> {path} '
          f'({os.path.getsize(path) / 1e6:.2f} MB)')


def main():
    if not ref_shim.available():
        raise SystemExit('/root/reference not present: fixtures can only be generated where the reference is mounted')
    model = S.make_smpl_model(0)
    model_t = S.smpl_model_to_torch(model)
    for name, cfg in CASES.items():
        run_case(name, cfg, model, model_t)


if __name__ == '__main__':
    main()
