"""TEST INFRASTRUCTURE.  How errors of the dense-volume gradients propagate into the sparse encoder's parameter gradients, measured on the REFERENCE alone (CPU):
re-runs the reference training step of oracle/gen_golden_training.py with the gradients of the three `.dense()` volumes multiplied by
(1 + eps * N(0,1)) elementwise and prints the relative change of the encoder gradients against the fixture.  usage: python
tests/helpers/encoder_grad_conditioning.py [eps=5e-5].  Result (eps = 5e-5): conv3.7.* 5e-6 .. 7e-6, conv3.6.weight 2.9e-5, conv3.1.bias 7.3e-5,
conv0.* 1.5e-4 .. 2.1e-4 -- the train() BatchNorms subtract per-channel means of gradients that are far from zero-mean here, so a
perturbation grows ~ 36 x from the last layer to the first.  tests/test_training_gpu.py sees the same profile (7.8e-5 -> 4.8e-3)."""
import sys, types, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import gen_golden_training as G, ref_shim, spconv_shim
from sherf_b200 import synthetic as S
import torch.nn as nn
model = S.make_smpl_model(0); model_t = S.smpl_model_to_torch(model)
cfg = G.CASES['training_step_32x32x16']
EPS = float(sys.argv[1]) if len(sys.argv) > 1 else 5e-5
orig_dense = spconv_shim.SparseConvTensor.dense
gen = torch.Generator().manual_seed(99)
def noisy_dense(self, channels_first=True):
    v = orig_dense(self, channels_first)
    if v.requires_grad and EPS > 0:
        v.register_hook(lambda g: g * (1 + EPS * torch.randn(g.shape, generator=gen)))
    return v
def run():
    spec = cfg['spec']
    scene = S.make_scene(spec, model); scene['rendering_options']['density_noise'] = 0
    tgt_img, tgt_mask = G.targets(spec, cfg['target_seed'])
    ren, dec, proj, state0 = G.initial_state(cfg, model_t)
    ref_tp = sys.modules['training.triplane']
    planes = scene['planes'].reshape(1, 96, 256, 256).clone().requires_grad_(True)
    feat = scene['obs_input_feature'].clone().requires_grad_(True)
    ren.train().requires_grad_(True); dec.train().requires_grad_(True); proj.requires_grad_(True)
    fake = types.SimpleNamespace(renderer=ren, decoder=dec, conv1d_projection=proj, use_3d_feature=True, neural_rendering_resolution=64, _last_planes=None,
        rendering_kwargs=dict(scene['rendering_options']), superresolution=None, encoder_2d_feature=lambda img, extract_feature=False: feat,
        backbone=types.SimpleNamespace(synthesis=lambda ws, update_emas=False, **k: planes))
    fake.prepare_sp_input = types.MethodType(ref_tp.TriPlaneGenerator.prepare_sp_input, fake)
    o_proj, nthreads = ren.projection, torch.get_num_threads()
    def p1(*a, **k):
        torch.set_num_threads(1)
        try: return o_proj(*a, **k)
        finally: torch.set_num_threads(nthreads)
    ren.projection = p1
    out = ref_tp.TriPlaneGenerator.synthesis(fake, None, scene['input_data'], None, use_sr_module=False, test_flag=False)
    G.the_loss(out, tgt_img, tgt_mask).backward()
    return {k: p.grad.clone() for k, p in ren.encoder_3d.named_parameters() if p.grad is not None}
g0 = np.load(__import__('os').path.join(G.OUT_DIR, 'training_step_32x32x16.npz'))
spconv_shim.SparseConvTensor.dense = noisy_dense
g1 = run()
print(f'dense-volume gradients perturbed by relative {EPS:g} (elementwise, random) -> change of the encoder gradients vs the fixture')
for k in ['conv3.7.bias','conv3.7.weight','conv3.6.weight','conv3.4.bias','conv3.3.weight','conv3.1.bias','down2.1.bias','conv2.7.bias','conv1.1.bias','conv0.4.bias','conv0.1.weight','conv0.0.weight']:
    key = 'g/renderer.encoder_3d.'+k; sub=False
    if key not in g0.files: key='gs/renderer.encoder_3d.'+k; sub=True
    ref = torch.from_numpy(g0[key]); mine = g1[k].reshape(-1)[::5] if sub else g1[k]
    print(f'  {k:16s} {float((mine-ref).norm()/ref.norm()):.2e}')
