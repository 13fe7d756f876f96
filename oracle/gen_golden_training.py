"""TEST INFRASTRUCTURE.  Golden vector of ONE TRAINING STEP at the generator boundary (SURVEY.md 8 f2 + f1): runs the REFERENCE's own
`TriPlaneGenerator.synthesis` (triplane.py:81-172) -> `ImportanceRenderer.forward` (renderer.py:286-437) -> `SparseConvNet.forward`
(:744-785, on the functional spconv stand-ins of oracle/spconv_shim.py) -> `NeRFDecoder.forward` (triplane.py:285-316) in train() mode on the
CPU, takes the reference's reconstruction loss (loss.py:150-151,167) and calls `loss.backward()` (loss.py:175).  Written to
tests/golden/training_step_<case>.npz: loss, image, the gradient of every parameter on the path (39 hot-path tensors, 39 sparse-encoder
tensors, conv1d_projection), of the tri-planes and of the 2-D feature map (leaf tensors standing in for the StyleGAN2 backbone / ResNet-18
encoder, SURVEY.md section 2: out of scope), and the BatchNorm running statistics after the step.

    python -m oracle.gen_golden_training          (a few minutes of CPU: dense conv3d on the 5 mm canonical grid)

The initial weights travel with the fixture where they are small (hot path 0.8 MB, projection conv); the 2.4 M sparse-encoder weights are
regenerated from their seed by the GPU test (tests/test_training_gpu.py); `state_sha256` over the whole initial state proves identity.
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
from oracle import ref_shim, sparse_encoder as SE

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
SUBSAMPLE_ABOVE, SUBSAMPLE_STRIDE = 100000, 5
VOL_ROW_STRIDE = 4
CASES = {'training_step_32x32x16': dict(spec=S.SceneSpec(H=32, W=32, samples=16, seed=17), weight_seed=5, enc_seed=6, proj_seed=7, target_seed=3)}


def targets(spec, seed):
    gen = torch.Generator().manual_seed(seed)
    return torch.rand(1, 3, spec.H, spec.W, generator=gen), (torch.rand(1, 1, spec.H, spec.W, generator=gen) > 0.4).float()


def the_loss(out, tgt_img, tgt_mask):
    """loss.py:150-151,167: 100 * mse(image / 2 + 0.5, target) + 10 * mse(weights_image, mask)."""
    return 100.0 * ((out['image'] / 2 + 0.5 - tgt_img) ** 2).mean() + 10.0 * ((out['weights_image'] - tgt_mask) ** 2).mean()


def initial_state(cfg, model_t):
    """(renderer, decoder, projection conv) of the reference with the fixture's seeded weights, and their flat state dict."""
    ren, dec = ref_shim.build_reference(model_t, cfg['weight_seed'])
    ref_renderer = sys.modules['training.volumetric_rendering.renderer']
    with torch.no_grad():
        dec.alpha_linear.weight *= 30                                   # an opaque body: densities that matter for the loss
        dec.alpha_linear.bias += 2.0
    torch.manual_seed(0)
    enc = ref_renderer.SparseConvNet(num_layers=4)                       # the reference's own class, on the functional spconv stand-ins
    enc.load_state_dict(SE.random_state_dict(enc, cfg['enc_seed']))
    ren.encoder_3d = enc
    torch.manual_seed(cfg['proj_seed'])
    proj = nn.Conv1d(96, 32, 1)
    state = {'renderer.' + k: v.clone() for k, v in ren.state_dict().items()}
    state.update({'decoder.' + k: v.clone() for k, v in dec.state_dict().items()})
    state.update({'conv1d_projection.' + k: v.clone() for k, v in proj.state_dict().items()})
    return ren, dec, proj, state


def state_checksum(state) -> str:
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(state[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def _synthesis_and_backward(ref_tp, fake, scene, tgt_img, tgt_mask):
    out = dict(ref_tp.TriPlaneGenerator.synthesis(fake, None, scene['input_data'], None, use_sr_module=False, test_flag=False))
    loss = the_loss(out, tgt_img, tgt_mask)
    loss.backward()
    out['loss'] = loss
    return out


def run_case(name, cfg, model, model_t):
    spec = cfg['spec']
    scene = S.make_scene(spec, model)
    scene['rendering_options']['density_noise'] = 0
    tgt_img, tgt_mask = targets(spec, cfg['target_seed'])
    ren, dec, proj, state0 = initial_state(cfg, model_t)
    ref_tp = sys.modules['training.triplane']
    planes = scene['planes'].reshape(1, 96, 256, 256).clone().requires_grad_(True)
    feat = scene['obs_input_feature'].clone().requires_grad_(True)
    ren.train().requires_grad_(True); dec.train().requires_grad_(True); proj.requires_grad_(True)
    fake = types.SimpleNamespace(
        renderer=ren, decoder=dec, conv1d_projection=proj, use_3d_feature=True, neural_rendering_resolution=64, _last_planes=None,
        rendering_kwargs=dict(scene['rendering_options']), superresolution=None,
        encoder_2d_feature=lambda img, extract_feature=False: feat,
        backbone=types.SimpleNamespace(synthesis=lambda ws, update_emas=False, **k: planes))
    fake.prepare_sp_input = types.MethodType(ref_tp.TriPlaneGenerator.prepare_sp_input, fake)
    # compute_normal's `norm[:, faces[:, k]] += n` (renderer.py:58-60) is an index_put WITHOUT accumulation over repeated vertex ids: the
    # sequential semantic (last face wins) is pinned by running THAT call single-threaded (see gen_golden_synthesis.py)
    o_proj, nthreads = ren.projection, torch.get_num_threads()

    def projection_single_thread(*a, **k):
        torch.set_num_threads(1)
        try:
            return o_proj(*a, **k)
        finally:
            torch.set_num_threads(nthreads)
    ren.projection = projection_single_thread
    # the gradients of the three `.dense()` volumes (renderer.py:762,771,780) at every VOL_ROW_STRIDE-th active voxel: what the render hands to
    # the sparse encoder's backward
    from oracle import spconv_shim
    taps, orig_dense = [], spconv_shim.SparseConvTensor.dense

    def tapped_dense(self, channels_first=True):
        v = orig_dense(self, channels_first)
        if v.requires_grad:
            keep = spconv_shim._first_rows(self.indices)[::VOL_ROW_STRIDE]
            idx = self.indices[keep][:, 1:]
            taps.append(idx)
            v.register_hook(lambda g_, idx=idx, slot=len(taps) - 1: taps.__setitem__(slot, (idx, g_[0][:, idx[:, 0], idx[:, 1], idx[:, 2]].t().clone())))
        return v
    spconv_shim.SparseConvTensor.dense = tapped_dense
    try:
        out = _synthesis_and_backward(ref_tp, fake, scene, tgt_img, tgt_mask)
    finally:
        spconv_shim.SparseConvTensor.dense = orig_dense
    loss = out.pop('loss')
    ren.projection = o_proj
    res = {
        'spec': np.array([spec.H, spec.W, spec.samples, spec.seed, int(spec.random_global_R), int(spec.white_back)], np.int64),
        'weight_seed': np.int64(cfg['weight_seed']), 'enc_seed': np.int64(cfg['enc_seed']), 'proj_seed': np.int64(cfg['proj_seed']),
        'target_seed': np.int64(cfg['target_seed']), 'state_sha256': np.array(state_checksum(state0)),
        'loss': np.float64(float(loss)), 'image': out['image'].detach().numpy(), 'weights_image': out['weights_image'].detach().numpy(),
    }
    for k, v in state0.items():
        if not k.startswith('renderer.encoder_3d.'):
            res['w/' + k] = v.numpy()
    n_grad = 0
    for prefix, mod in (('renderer.', ren), ('decoder.', dec), ('conv1d_projection.', proj)):
        for k, p in mod.named_parameters():
            if p.grad is not None:
                if p.grad.numel() > SUBSAMPLE_ABOVE:                      # the large sparse-conv weight gradients travel as every 5th element
                    res['gs/' + prefix + k] = p.grad.reshape(-1)[::SUBSAMPLE_STRIDE].numpy().copy()
                else:
                    res['g/' + prefix + k] = p.grad.numpy()
                n_grad += 1
            else:
                res['nograd/' + prefix + k] = np.zeros(0, np.float32)
    res['g/planes'], res['g/obs_input_feature'] = planes.grad.numpy(), feat.grad.numpy()
    assert len(taps) == 3 and all(isinstance(t, tuple) for t in taps), 'expected the gradients of three dense levels'
    for l, (idx, gv) in enumerate(taps):
        res[f'gvol{l}/zyx'], res[f'gvol{l}/g'] = idx.numpy().astype(np.int32), gv.numpy()
    for k, v in ren.state_dict().items():
        if 'running_' in k or 'num_batches_tracked' in k:
            res['stat/renderer.' + k] = v.numpy()
    path = os.path.join(OUT_DIR, name + '.npz')
    np.savez_compressed(path, **res)
    print(f'{name}: loss {float(loss):.6f}, {n_grad} parameter gradients + planes + feature map, acc.max={float(out["weights_image"].max()):.3f} -> {path} '
          f'({os.path.getsize(path) / 1e6:.2f} MB)')


def main():
    if not ref_shim.available():
        raise SystemExit('reference files not present: fixtures can only be generated where the reference is readable')
    model = S.make_smpl_model(0)
    model_t = S.smpl_model_to_torch(model)
    for name, cfg in CASES.items():
        run_case(name, cfg, model, model_t)


if __name__ == '__main__':
    main()
