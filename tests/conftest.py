import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sherf_b200 import synthetic as S  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_CASES = ['c1_64x64x16', 'ragged_45x38x24_R_white', 'default_init_32x32x48']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def smpl_model():
    return S.make_smpl_model(0)


@pytest.fixture(scope='session')
def smpl_model_t(smpl_model):
    return S.smpl_model_to_torch(smpl_model)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    g = {k: z[k] for k in z.files}
    H, W, samples, seed, rgr, wb = [int(v) for v in g['spec']]
    g['scene_spec'] = S.SceneSpec(H=H, W=W, samples=samples, seed=seed, random_global_R=bool(rgr), white_back=bool(wb))
    g['weights'] = {k[2:]: torch.from_numpy(g[k]) for k in g if k.startswith('w/')}
    return g


def scene_to(scene, device):
    """Deep-copy a synthetic scene's tensors to `device`."""
    def mv(x):
        if torch.is_tensor(x):
            return x.to(device)
        if isinstance(x, dict):
            return {k: mv(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(mv(v) for v in x)
        return x
    return {k: mv(v) for k, v in scene.items()}


def modules_from_weights(weights, smpl_model, mlp_precision='bf16x3'):
    """sherf_b200 modules carrying the given hot-path state dict ('renderer.*' / 'decoder.*' names)."""
    from sherf_b200.triplane import hot_path_modules
    ren, dec = hot_path_modules(smpl_model, seed=0, mlp_precision=mlp_precision)
    rsd = {k[len('renderer.'):]: v for k, v in weights.items() if k.startswith('renderer.')}
    dsd = {k[len('decoder.'):]: v for k, v in weights.items() if k.startswith('decoder.')}
    missing, unexpected = ren.load_state_dict(rsd, strict=False)
    assert not unexpected, unexpected
    assert all(m.startswith('encoder_3d') for m in missing), missing
    dec.load_state_dict(dsd, strict=True)
    return ren, dec
