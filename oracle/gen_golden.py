"""TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the REFERENCE's own ImportanceRenderer.forward +
NeRFDecoder (imported under the shims of oracle/ref_shim.py) on seeded synthetic scenes, in THIS container
(/root/reference must exist).  The fixtures travel to the GPU box; the reference cannot.

    python -m oracle.gen_golden            # rewrites tests/golden/

Each fixture stores: the scene spec (regenerated from the seed by sherf_b200.synthetic, with input checksums to
prove regeneration is bit-identical), the hot-path weights the reference modules were initialised with, the three
forward outputs, and stage-wise taps captured by wrapping (not editing) reference methods.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

from sherf_b200 import synthetic as S
from oracle import ref_shim, port

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = {
    # BASELINE.json configs[0]: 64x64 render, 16 samples/ray
    'c1_64x64x16': dict(spec=S.SceneSpec(H=64, W=64, samples=16, seed=0), weight_seed=0, dense_sigma=True),
    # ragged sizes (N % 8 != 0, S % 32 != 0), global rotation R != I (HuMMan / ZJU style), white background
    'ragged_45x38x24_R_white': dict(spec=S.SceneSpec(H=38, W=45, samples=24, seed=3, random_global_R=True, white_back=True),
                                    weight_seed=1, dense_sigma=True),
    # PyTorch default init untouched (tiny densities)
    'default_init_32x32x48': dict(spec=S.SceneSpec(H=32, W=32, samples=48, seed=5), weight_seed=2, dense_sigma=False),
}


def checksum(scene) -> str:
    h = hashlib.sha256()
    for k in ('planes', 'obs_input_img', 'obs_input_feature', 'ray_origins', 'ray_directions', 'near', 'far'):
        h.update(scene[k].numpy().tobytes())
    for v in scene['volumes']:
        h.update(v.numpy().tobytes())
    idt = scene['input_data']
    for k in ('vertices', 't_vertices', 'obs_vertices', 't_world_bounds', 'obs_K_all', 'obs_R_all', 'obs_T_all'):
        h.update(idt[k].numpy().tobytes())
    for pk in ('params', 't_params', 'obs_params'):
        for k in ('poses', 'shapes', 'R', 'Th'):
            h.update(idt[pk][k].numpy().tobytes())
    return h.hexdigest()


def run_case(name, cfg, model, model_t):
    ren, dec = ref_shim.build_reference(model_t, cfg['weight_seed'])
    if cfg['dense_sigma']:
        with torch.no_grad():
            dec.alpha_linear.weight *= 30
            dec.alpha_linear.bias += 2.0
    scene = S.make_scene(cfg['spec'], model)
    cap = {'knn': []}
    ref_renderer = sys.modules['training.volumetric_rendering.renderer']
    orig_knn = ref_renderer.knn_points

    def knn_tap(*a, **k):
        r = orig_knn(*a, **k)
        cap['knn'].append((r[0][0, :, 0].clone(), r[1][0, :, 0].clone()))
        return r
    ref_renderer.knn_points = knn_tap
    o_t2c, o_c2s, o_proj, o_run = ren.coarse_deform_target2c, ren.coarse_deform_c2source, ren.projection, ren.run_model

    def t2c(*a, **k):
        r = o_t2c(*a, **k); cap['can'], cap['cdir'] = r[0][0], r[1][0]; return r

    def c2s(*a, **k):
        r = o_c2s(*a, **k); cap['world'] = r[1][0]; return r

    def proj(*a, **k):
        r = o_proj(*a, **k); cap['uv'] = r[0, 0]; return r

    def run_model(planes, f2d, f3d, *a, **k):
        cap['f2d'], cap['f3d'] = f2d[0], f3d[0]; return o_run(planes, f2d, f3d, *a, **k)
    ren.coarse_deform_target2c, ren.coarse_deform_c2source, ren.projection, ren.run_model = t2c, c2s, proj, run_model
    orig_sfp = ref_renderer.sample_from_planes

    def sfp_tap(*a, **k):
        r = orig_sfp(*a, **k); cap['tri'] = r[0]; return r          # [3, P, 32]  (renderer.py:234-243, called at :402)
    ref_renderer.sample_from_planes = sfp_tap
    hk = dec.register_forward_hook(lambda m, i, o: cap.update(sigma=o['sigma'][0, :, 0], rgbp=o['rgb'][0], tok=i[1]))
    h3 = ren.encoder_3d.register_forward_hook(lambda m, i, o: cap.update(f3raw=o[0]))
    try:
        rgb, depth, acc = ref_shim.render(ren, dec, scene)
    finally:
        ref_renderer.knn_points = orig_knn
        ref_renderer.sample_from_planes = orig_sfp
        hk.remove(); h3.remove()
    d2_1, id_1 = cap['knn'][0]
    mask = d2_1 < (0.05 ** 2)
    sel = mask.nonzero()[:, 0]
    P = sel.numel()
    k = min(P, 256)
    w = port.hot_path_state_dict(ren, dec)
    spec = cfg['spec']
    out = {
        'spec': np.array([spec.H, spec.W, spec.samples, spec.seed, int(spec.random_global_R), int(spec.white_back)], np.int64),
        'weight_seed': np.int64(cfg['weight_seed']), 'input_sha256': np.array(checksum(scene)),
        'rgb': rgb[0].numpy(), 'depth': depth[0].numpy(), 'acc': acc[0].numpy(),
        'mask_bits': np.packbits(mask.numpy()), 'num_points': np.int64(P),
        'id1': id_1[sel].numpy().astype(np.int16), 'id3': cap['knn'][2][1].numpy().astype(np.int16),
        'can': cap['can'].numpy(), 'cdir': cap['cdir'].numpy(), 'uv': cap['uv'].numpy(),
        'sigma': cap['sigma'].numpy(), 'rgb_pts': cap['rgbp'].numpy(),
        'tok01_head': cap['tok'][:2, :k].permute(1, 0, 2).reshape(k, 64).numpy(),
        'f2d_head': cap['f2d'][:k].numpy(), 'f3raw_head': cap['f3raw'][:k].numpy(),
        'tri_head': cap['tri'][:, :k].permute(1, 0, 2).reshape(k, 96).numpy(),
    }
    for name_w, t in w.items():
        out['w/' + name_w] = t.numpy()
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: N={rgb.shape[1]} S={spec.samples} P={P} acc.max={float(acc.max()):.3f} -> {path} '
          f'({os.path.getsize(path) / 1e6:.2f} MB)')


IMPORTANCE_CASES = {
    # SURVEY a13 / BASELINE configs[4] in miniature: coarse + fine importance samples through the REPAIRED fine pass
    # (oracle/ref_shim.render_importance: the reference's own sample_importance / unify_samples / ray marcher)
    'importance_28x20x16p12': dict(spec=S.SceneSpec(H=20, W=28, samples=16, seed=9, random_global_R=True), n_importance=12,
                                   weight_seed=4, u_seed=3),
}


def importance_u(n_rays, n_importance, seed):
    """The uniform draws standing for torch.rand at renderer.py:526 (CPU generator: identical on every machine)."""
    return torch.rand(n_rays, n_importance, generator=torch.Generator().manual_seed(seed))


def run_importance_case(name, cfg, model, model_t):
    ren, dec = ref_shim.build_reference(model_t, cfg['weight_seed'])
    with torch.no_grad():
        dec.alpha_linear.weight *= 30
        dec.alpha_linear.bias += 2.0
    scene = S.make_scene(cfg['spec'], model)
    scene['rendering_options']['depth_resolution_importance'] = cfg['n_importance']
    spec = cfg['spec']
    u = importance_u(spec.H * spec.W, cfg['n_importance'], cfg['u_seed'])
    rgb, depth, acc, st = ref_shim.render_importance(ren, dec, scene, u, return_stages=True)
    out = {
        'spec': np.array([spec.H, spec.W, spec.samples, spec.seed, int(spec.random_global_R), int(spec.white_back)], np.int64),
        'n_importance': np.int64(cfg['n_importance']), 'u_seed': np.int64(cfg['u_seed']),
        'weight_seed': np.int64(cfg['weight_seed']), 'input_sha256': np.array(checksum(scene)),
        'rgb': rgb[0].numpy(), 'depth': depth[0].numpy(), 'acc': acc[0].numpy(),
        'coarse_weights': st['coarse_weights'].numpy(), 't_fine': st['t_fine'].numpy(),
        'sigma_coarse': st['sigma_coarse'].numpy(), 'sigma_fine': st['sigma_fine'].numpy(), 'colors_fine': st['colors_fine'].numpy(),
    }
    for name_w, t in port.hot_path_state_dict(ren, dec).items():
        out['w/' + name_w] = t.numpy()
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: N={rgb.shape[1]} S={spec.samples}+{cfg["n_importance"]} fine survivors={int((st["sigma_fine"] != -80).sum())} '
          f'acc.max={float(acc.max()):.3f} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)')


def main():
    if not ref_shim.available():
        raise SystemExit('/root/reference not present: fixtures can only be generated where the reference is mounted')
    model = S.make_smpl_model(0)
    model_t = S.smpl_model_to_torch(model)
    only = sys.argv[1:]
    for name, cfg in CASES.items():
        if not only or name in only:
            run_case(name, cfg, model, model_t)
    for name, cfg in IMPORTANCE_CASES.items():
        if not only or name in only:
            run_importance_case(name, cfg, model, model_t)


if __name__ == '__main__':
    main()
