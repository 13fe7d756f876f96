"""One full training step at the generator boundary on the GPU box, BASELINE configs[1] size (512x512 rays x 64 samples): the overlay
`TriPlaneGenerator.synthesis` in train() (vertex features -> sparse 3-D encoder with batch-statistics BatchNorm -> render) + the reference's
reconstruction loss (loss.py:150-151,167) + loss.backward() into every parameter on the path, the tri-planes and the 2-D feature map (leaf
tensors stand in for the StyleGAN2 backbone / ResNet-18 encoder: SURVEY.md section 2, out of scope).  Prints one JSON line (CUDA events)."""
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sherf_b200 import overlay, synthetic as S                      # noqa: E402


class LeafPlanes(nn.Module):
    def __init__(self, planes):
        super().__init__()
        self.planes = nn.Parameter(planes.reshape(1, 96, 256, 256).clone())

    def mapping(self, z, c, **k):
        return None

    def synthesis(self, ws, update_emas=False, **k):
        return self.planes


class LeafFeature(nn.Module):
    def __init__(self, feat):
        super().__init__()
        self.feat = nn.Parameter(feat.clone())

    def forward(self, x, extract_feature=False):
        return self.feat if extract_feature else x.new_zeros(x.shape[0], 512)


def main():
    H = W = int(os.environ.get('BWD_RES', '512'))
    samples = int(os.environ.get('BWD_SAMPLES', '64'))
    steps, warmup = int(os.environ.get('BWD_STEPS', '5')), int(os.environ.get('BWD_WARMUP', '2'))
    dev = torch.device('cuda:0')
    model = S.make_smpl_model(0)
    cpu_scene = S.make_scene(S.SceneSpec(H=H, W=W, samples=samples, seed=0), model)
    cpu_scene['rendering_options']['density_noise'] = 0

    def mv(x):
        if torch.is_tensor(x):
            return x.to(dev)
        if isinstance(x, dict):
            return {k: mv(v) for k, v in x.items()}
        if isinstance(x, list):
            return [mv(v) for v in x]
        return x
    scene = {k: mv(v) for k, v in cpu_scene.items()}
    overlay.install()
    try:
        overlay.set_factories(backbone=lambda *a, **k: LeafPlanes(cpu_scene['planes']), encoder_2d=lambda: LeafFeature(cpu_scene['obs_input_feature']))
        rendering = dict(cpu_scene['rendering_options'], c_gen_conditioning_zero=True, superresolution_noise_mode='none')
        cwd = os.getcwd()
        os.chdir('/tmp')
        try:
            G = overlay.construct_class_by_name(class_name='training.triplane.TriPlaneGenerator', z_dim=512, c_dim=0, w_dim=512, use_1d_feature=True,
                                                use_2d_feature=True, use_3d_feature=True, use_trans=True, use_NeRF_decoder=True, img_resolution=512,
                                                img_channels=3, rendering_kwargs=rendering)
        finally:
            os.chdir(cwd)
    finally:
        overlay.set_factories()
        overlay.uninstall()
    G.renderer.set_smpl_model(model)
    G = G.to(dev).train().requires_grad_(True)
    tgt = torch.rand(1, 3, H, W, device=dev)
    mask = (torch.rand(1, 1, H, W, device=dev) > 0.4).float()

    def step():
        for p in G.parameters():
            p.grad = None
        out = G.synthesis(None, scene['input_data'], None, use_sr_module=False, test_flag=False)
        loss = 100.0 * ((out['image'] / 2 + 0.5 - tgt) ** 2).mean() + 10.0 * ((out['weights_image'] - mask) ** 2).mean()
        loss.backward()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n_grads = sum(p.grad is not None for p in G.parameters())
    print(json.dumps({'what': 'generator-level training step: synthesis (vertex features, sparse encoder train(), render) + loss + backward', 'H': H, 'W': W,
                      'samples': samples, 'surviving_points': G.renderer.last_num_points, 'ms_per_step': ms,
                      'ray_samples_per_sec_training': H * W * samples / ms * 1e3, 'gradient_tensors': n_grads, 'loss': float(loss),
                      'peak_mem_GB': torch.cuda.max_memory_allocated() / 2**30}))


if __name__ == '__main__':
    main()
