"""On-disk formats of the reference: `network-snapshot-*.pkl` in, SHERF generator with the CUDA render path out.

The reference writes whole-module pickles every tick (training_loop.py:563-579: `dict(G=..., G_ema=..., ...)`, classes decorated with
`torch_utils.persistence` so the pickle carries the SOURCE of `training/triplane.py`) and resumes by constructing the generator by
name and copying tensors by name (training_loop.py:193-208).  `load_generator` does exactly that with the overlay installed:
  1. `legacy.load_network_pkl` (legacy.py:24-61) unpickles the snapshot.  The pickled-source generator executes the reference's own
     triplane.py, whose `training.volumetric_rendering.renderer` import resolves to sherf_b200 through the overlay, and whose spconv /
     imageio imports resolve to parameter-container stubs where those packages are absent;
  2. `dnnlib.util.construct_class_by_name('training.triplane.TriPlaneGenerator', **snapshot.init_kwargs)` builds the overlay generator;
  3. `misc.copy_params_and_buffers(snapshot, G, require_all=True)` moves all tensors by name (563 for the shipped configuration);
  4. the SMPL model the snapshot's renderer carries (renderer.py:283-284) is handed to the new renderer.
Needs the reference tree (dnnlib, torch_utils, legacy, the backbone / encoder classes) on `reference_root`; nothing of it is copied.
"""
from __future__ import annotations

import os
import sys

import torch


def load_generator(pkl_path: str, reference_root: str, which: str = 'G_ema', device=None):
    from . import overlay
    overlay.install()
    reference_root = os.path.abspath(reference_root)
    if not os.path.isfile(os.path.join(reference_root, 'legacy.py')):
        raise FileNotFoundError(f'{reference_root} is not the reference\'s `sherf/` directory (legacy.py not found)')
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import dnnlib
    import legacy
    from torch_utils import misc
    with open(pkl_path, 'rb') as f:
        data = legacy.load_network_pkl(f)
    src = data[which]
    kwargs = dict(src.init_kwargs)
    cwd = os.getcwd()
    try:
        if not os.path.exists(os.path.join('assets', 'SMPL_NEUTRAL.pkl')):
            os.chdir(reference_root)                                  # the constructor looks for assets/SMPL_NEUTRAL.pkl like renderer.py:283
        G = dnnlib.util.construct_class_by_name(class_name='training.triplane.TriPlaneGenerator', **kwargs)
    finally:
        os.chdir(cwd)
    with torch.no_grad():
        misc.copy_params_and_buffers(src, G, require_all=True)
    smpl = getattr(getattr(src, 'renderer', None), 'SMPL_NEUTRAL', None)
    if isinstance(smpl, dict) and G.renderer.SMPL_NEUTRAL is None:
        G.renderer.set_smpl_model({k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in smpl.items()})
    G = G.eval().requires_grad_(False)
    return G.to(device) if device is not None else G
