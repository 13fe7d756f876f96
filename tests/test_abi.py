"""CPU tests of the drop-in boundary: the C-ABI library loads here (no GPU), exports every symbol the header declares,
the ctypes mirrors match the C struct layouts, and the product modules refuse to run without CUDA."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT
from sherf_b200 import _lib, build

HEADER = os.path.join(ROOT, 'include', 'sherf_b200.h')


@pytest.fixture(scope='module')
def lib():
    build.build_library()
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(sherf_[a-z_0-9]+)\s*\(', src)))


def test_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert set(names) == set(_lib.EXPORTS), (names, _lib.EXPORTS)
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'
    assert lib.sherf_abi_version() == 5


def test_struct_layouts_match_header(tmp_path):
    structs = ['SherfSmplModel', 'SherfPose', 'SherfFrame', 'SherfScene', 'SherfWeights', 'SherfRays', 'SherfOptions', 'SherfOut',
               'SherfDebug', 'SherfSparseConv', 'SherfSparseEncoder', 'SherfObservation', 'SherfOutGrads', 'SherfWeightGrads',
               'SherfInputGrads']
    prog = '#include <stdio.h>\n#include "sherf_b200.h"\nint main(){' + ''.join(
        f'printf("{s} %zu\\n", sizeof({s}));' for s in structs) + 'return 0;}'
    c = tmp_path / 'sz.c'
    c.write_text(prog)
    exe = tmp_path / 'sz'
    subprocess.run(['gcc', '-I', os.path.dirname(HEADER), str(c), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    sizes = dict(l.split() for l in out.strip().splitlines())
    for s in structs:
        assert int(sizes[s]) == ctypes.sizeof(getattr(_lib, s)), s


def test_scratch_bytes_and_argument_validation(lib):
    sc = _lib.SherfScene()
    sc.plane_ch, sc.plane_h, sc.plane_w = 32, 256, 256
    sc.img_h = sc.img_w = 64
    sc.feat_ch, sc.feat_h, sc.feat_w = 64, 32, 32
    for l, (c, d) in enumerate(zip((32, 64, 96), ((48, 176, 192), (24, 88, 96), (12, 44, 48)))):
        sc.vol_ch[l] = c
        for a in range(3):
            sc.vol_dim[l][a] = d[a]
    small = lib.sherf_scratch_bytes(ctypes.byref(sc), 4096, 16, 0, 6890)
    big = lib.sherf_scratch_bytes(ctypes.byref(sc), 512 * 512, 64, 0, 6890)
    assert 0 < small < big < 8 << 30
    assert lib.sherf_scratch_bytes(ctypes.byref(sc), 0, 16, 0, 6890) == 0
    rc = lib.sherf_render_forward(None, None, None, None, None, None, None, None, None, 0, None, None)
    assert rc == -1 and b'null' in lib.sherf_last_error()
    # backward arena = forward arena + its own buffers; argument checks run before any CUDA call
    bsmall = lib.sherf_backward_scratch_bytes(ctypes.byref(sc), 4096, 16, 6890)
    bbig = lib.sherf_backward_scratch_bytes(ctypes.byref(sc), 512 * 512, 64, 6890)
    assert small < bsmall < bbig < 40 << 30          # one backward chunk of up to 2^20 points keeps 24 GB of activations
    assert lib.sherf_backward_scratch_bytes(ctypes.byref(sc), 0, 16, 6890) == 0
    assert lib.sherf_render_backward(None, None, None, None, None, None, None, None, None, None, 0, None, None) == -1
    assert ctypes.sizeof(_lib.SherfWeightGrads) == ctypes.sizeof(_lib.SherfWeights) == 39 * ctypes.sizeof(ctypes.c_void_p)


def test_no_cpu_fallback(smpl_model):
    from sherf_b200 import synthetic as S
    from sherf_b200.triplane import hot_path_modules
    ren, dec = hot_path_modules(smpl_model)
    scene = S.make_scene(S.SceneSpec(H=4, W=4, samples=4), smpl_model)
    with pytest.raises(RuntimeError, match='CUDA'):
        ren(scene['planes'], scene['obs_input_img'], scene['obs_input_feature'], scene['volumes'], None, scene['obs_sp_input'], dec,
            scene['ray_origins'], scene['ray_directions'], scene['near'], scene['far'], scene['input_data'], scene['rendering_options'])
    with pytest.raises(RuntimeError):
        dec(torch.zeros(1, 39), torch.zeros(3, 1, 32), torch.zeros(1, 27))


def test_checkpoint_names_match_reference_contract(smpl_model):
    """resume copies parameters by name with require_all=True."""
    from sherf_b200.triplane import hot_path_modules
    ren, dec = hot_path_modules(smpl_model)
    names = set(ren.state_dict().keys())
    for n in ['conv1d_projection.weight', 'conv1d_projection.bias', 'conv1d_reprojection.weight', 'conv1d_reprojection.bias',
              'transformer.layers.0.0.fn.norm.weight', 'transformer.layers.0.0.fn.fn.to_qkv.weight',
              'transformer.layers.0.0.fn.fn.to_out.0.weight', 'transformer.layers.0.0.fn.fn.to_out.0.bias',
              'transformer.layers.0.1.fn.norm.bias', 'transformer.layers.0.1.fn.fn.net.0.weight',
              'transformer.layers.0.1.fn.fn.net.3.bias', 'rgb_enc._freqs', 'pos_enc._phases', 'view_enc._freqs',
              'encoder_3d.conv0.0.weight', 'encoder_3d.conv0.1.running_mean', 'encoder_3d.down3.0.weight', 'encoder_3d.conv4.6.weight']:
        assert n in names, n
    assert ren.state_dict()['conv1d_projection.weight'].shape == (96, 192, 1)
    dn = dec.state_dict()
    assert dn['pts_linears.0.weight'].shape == (128, 71) and dn['pts_linears.5.weight'].shape == (128, 199)
    assert dn['views_linear.weight'].shape == (64, 187) and dn['alpha_linear.weight'].shape == (1, 128)
    n_hot = sum(v.numel() for k, v in ren.state_dict().items() if k.startswith(('conv1d', 'transformer'))) + sum(
        p.numel() for p in dec.parameters())
    assert n_hot == 192804                                                       # SURVEY.md section 7 "hard parts"


def test_importance_and_sparse_argument_validation(lib):
    """Argument checks run before any CUDA call, so they can be exercised without a GPU: the importance pass needs S >= 3 and the
    uniform draws; the sparse encoder and the ray generator reject null / degenerate arguments with SHERF_E_INVALID."""
    sc = _lib.SherfScene()
    sc.plane_ch, sc.plane_h, sc.plane_w = 32, 8, 8
    sc.img_h = sc.img_w = 8
    sc.feat_ch, sc.feat_h, sc.feat_w = 64, 4, 4
    for l, c in enumerate((32, 64, 96)):
        sc.vol_ch[l] = c
        for a in range(3):
            sc.vol_dim[l][a] = 4
    with_f = lib.sherf_scratch_bytes(ctypes.byref(sc), 1024, 16, 16, 6890)
    without = lib.sherf_scratch_bytes(ctypes.byref(sc), 1024, 16, 0, 6890)
    assert with_f > without > 0                                    # the fine pass carries its own per-sample bookkeeping
    assert lib.sherf_scratch_bytes(ctypes.byref(sc), 1024, 16, -1, 6890) == 0
    one = ctypes.c_float(0)
    p = ctypes.addressof(one)                                       # any non-null address: validation never dereferences it
    smpl, fr, w, rays, opts, out = _lib.SherfSmplModel(), _lib.SherfFrame(), _lib.SherfWeights(), _lib.SherfRays(), _lib.SherfOptions(), _lib.SherfOut()
    smpl.weights = smpl.posedirs = p
    smpl.n_verts = 6890
    sc.planes = sc.obs_img = sc.obs_feat = p
    for l in range(3):
        sc.vol[l] = p
    rays.origins = rays.dirs = rays.near_ = rays.far_ = p
    out.rgb = out.depth = out.acc = p
    rays.n_rays, rays.n_samples, rays.n_importance = 16, 2, 4
    args = [ctypes.byref(x) for x in (smpl, fr, sc, w, rays, opts, out)] + [None, None, 0, None, None]
    assert lib.sherf_render_forward(*args) == -1 and b'n_samples >= 3' in lib.sherf_last_error()
    rays.n_samples = 8
    assert lib.sherf_render_forward(*args) == -1 and b'importance_u' in lib.sherf_last_error()
    rays.n_importance = 300
    assert lib.sherf_render_forward(*args) == -1 and b'n_importance' in lib.sherf_last_error()
    sh = (ctypes.c_int32 * 3)(32, 64, 64)
    assert lib.sherf_sparse_encoder_scratch_bytes(100, sh) > 4 * 32 * 64 * 64
    assert lib.sherf_sparse_encoder_scratch_bytes(0, sh) == 0
    assert lib.sherf_sparse_encode(None, None, None, 0, None, None, None, None, None, 0, None) == -1
    assert lib.sherf_generate_rays(None, None, None, 4, 4, None, None, None, None, None, None, None) == -1
    assert lib.sherf_debug_sample_importance(None, None, None, None, None, None) == -1


def test_cuda_only_entry_points_refuse_cpu():
    from sherf_b200.rays import generate_rays
    from sherf_b200.renderer import SparseConvNet, SparseConvTensor
    import numpy as np
    with pytest.raises(RuntimeError, match='CUDA'):
        generate_rays(4, 4, np.eye(3), np.eye(3), np.zeros(3), np.array([[0, 0, 0], [1, 1, 1.0]]), 'cpu')
    enc = SparseConvNet(4).eval()
    sp = SparseConvTensor(torch.zeros(3, 32), torch.zeros(3, 4, dtype=torch.int32), [32, 32, 32], 1)
    with pytest.raises(RuntimeError, match='CUDA'):
        enc(sp)
    with pytest.raises(RuntimeError, match='CUDA'):                    # train() exists since ABI v5, on CUDA tensors only as well
        SparseConvNet(4).train()(sp)


def test_plain_c_client_binds_the_library(tmp_path, lib):
    """include/sherf_b200.h compiles as C99 and a dlopen()ing C program resolves and calls the entry points (tests/c/cabi_client.c)."""
    exe = tmp_path / 'cabi_client'
    src = os.path.join(ROOT, 'tests', 'c', 'cabi_client.c')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.dirname(HEADER), src, '-o', str(exe), '-ldl'], check=True)
    r = subprocess.run([str(exe), _lib.lib_path()], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith('ok abi=5')
    print(r.stdout.strip())


def test_modules_deepcopy_and_pickle_without_runtime_state(smpl_model):
    """training_loop.py:196,572-579: the generator is deep-copied and pickled every tick.  Device-side state (ctypes structs with raw
    pointers, the scratch arena) lives outside the modules' __dict__ (renderer._RUNTIME), so it can never leak into a copy."""
    import copy
    import pickle
    from sherf_b200 import renderer as R
    from sherf_b200.triplane import hot_path_modules
    ren, dec = hot_path_modules(smpl_model)
    rt = R._runtime(ren)
    rt.smpl_dev = ('cuda:0', [], _lib.SherfSmplModel())                      # what a forward leaves behind
    rt.w_cache = ((), _lib.SherfWeights(), [], [])
    rt.scratch = torch.empty(16, dtype=torch.uint8)
    with pytest.raises(ValueError):
        pickle.dumps(rt.w_cache[1])                                          # the reason they must stay out of the module
    ren2 = copy.deepcopy(ren)
    ren3 = pickle.loads(pickle.dumps(ren))
    for r in (ren2, ren3):
        assert R._runtime(r).scratch is None and R._runtime(r).w_cache is None and R._runtime(r).smpl_dev is None
        assert set(r.state_dict()) == set(ren.state_dict())
        assert all(torch.equal(v, r.state_dict()[k]) for k, v in ren.state_dict().items())
    assert not any(isinstance(v, ctypes.Structure) for v in vars(ren).values())
    ren.invalidate_weights()
    assert R._runtime(ren).w_cache is None
