"""The drop-in boundary as the reference's own machinery sees it: `training.triplane.TriPlaneGenerator` resolved to the
overlay, constructed by `dnnlib.util.construct_class_by_name`, resumed with `misc.copy_params_and_buffers(require_all=True)`, deep-copied
and pickled with `torch_utils.persistence` -- against the reference's OWN generator class built in the same interpreter.  CPU; needs
/root/reference (skipped on the GPU box, where tests/test_overlay_gpu.py covers the render half)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import ref_shim


@pytest.mark.skipif(not ref_shim.mounted(), reason='the full reference tree (backbone, super-resolution, dnnlib) is not mounted here')
def test_overlay_generator_matches_reference_names_and_resumes():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'helpers', 'overlay_vs_reference.py')], capture_output=True, text=True,
                       timeout=1200, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')][-1]
    out = json.loads(line[len('RESULT '):])
    print(out)
    assert out['ref_file'].startswith('/root/reference') and 'sherf_b200/overlay/triplane.py' in out['our_file']
    assert out['only_ref'] == [] and out['only_ours'] == [] and out['shape_mismatch'] == []
    assert out['state_dict_equal'] and out['copied_equal']
    assert out['pickle_roundtrip_names_equal'] and out['pickle_roundtrip_values_equal'] and out['init_kwargs_kept']
    assert out['hot_path_params'] == 192804


def test_overlay_finder_standalone_names():
    """Without the reference tree the dotted names still import (namespace stand-ins) and export the reference's import surface."""
    code = ('import sys; sys.path.insert(0, %r)\n'
            'from sherf_b200 import overlay\n'
            'overlay.install()\n'
            'import training.volumetric_rendering.renderer as r, training.triplane as t\n'
            'assert all(hasattr(r, n) for n in ("ImportanceRenderer", "read_pickle", "SMPL_to_tensor"))\n'
            'assert all(hasattr(t, n) for n in ("TriPlaneGenerator", "OSGDecoder", "NeRFDecoder", "ResNet18Classifier"))\n'
            'import spconv.pytorch as sp\n'
            'assert issubclass(sp.SubMConv3d, __import__("torch").nn.Module)\n'
            'cls = overlay.construct_class_by_name(32, class_name="training.triplane.NeRFDecoder")\n'
            'print("OK", type(cls).__name__)\n') % ROOT
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600, cwd='/tmp')
    assert r.returncode == 0 and 'OK NeRFDecoder' in r.stdout, r.stderr[-3000:]


@pytest.mark.skipif(not ref_shim.mounted(), reason='the full reference tree (legacy.py, dnnlib, backbone) is not mounted here')
def test_load_reference_snapshot_into_overlay_generator(tmp_path):
    """f4: a `network-snapshot-*.pkl` written from the REFERENCE's own generator (persistence pickle with the reference's triplane.py
    source, spconv classes referenced by their real module paths) -> sherf_b200.checkpoint.load_generator -> overlay generator carrying
    every tensor (legacy.load_network_pkl + construct_class_by_name + copy_params_and_buffers(require_all=True), training_loop.py:199-208),
    in a process that has neither spconv nor imageio."""
    pkl, sd = str(tmp_path / 'network-snapshot-000000.pkl'), str(tmp_path / 'state.pt')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'helpers', 'make_reference_snapshot.py'), pkl, sd], capture_output=True, text=True,
                       timeout=1200, cwd=ROOT)
    assert r.returncode == 0 and 'SNAPSHOT_OK' in r.stdout, r.stderr[-3000:]
    code = ('import sys, torch; sys.path.insert(0, %r)\n'
            'from sherf_b200.checkpoint import load_generator\n'
            'G = load_generator(%r, "/root/reference/sherf")\n'
            'want = torch.load(%r)\n'
            'got = G.state_dict()\n'
            'assert set(got) == set(want), (sorted(set(got) ^ set(want))[:5])\n'
            'assert all(torch.equal(got[k], want[k]) for k in want)\n'
            'assert "sherf_b200/overlay/triplane.py" in sys.modules["training.triplane"].__file__ and isinstance(G, sys.modules["training.triplane"].TriPlaneGenerator)\n'
            'assert type(G.renderer).__module__ == "sherf_b200.renderer" and G.renderer.SMPL_NEUTRAL is not None\n'
            'assert G.rendering_kwargs["depth_resolution"] == 48 and not G.training\n'
            'print("LOAD_OK", len(got))\n') % (ROOT, pkl, sd)
    import torchvision  # noqa: F401
    r = subprocess.run([sys.executable, '-c', 'import torchvision.models as t; o = t.resnet18; t.resnet18 = lambda *a, pretrained=False, **k: o(weights=None)\n' + code],
                       capture_output=True, text=True, timeout=1200, cwd=str(tmp_path))
    assert r.returncode == 0 and 'LOAD_OK 563' in r.stdout, (r.stdout[-500:], r.stderr[-3000:])
