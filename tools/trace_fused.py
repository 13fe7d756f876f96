"""Cycle-counter breakdown of the fused decoder kernels (needs a library built with SHERF_FUSED_TRACE=1).

usage: SHERF_FUSED_TRACE=1 python sherf_b200/build.py --force && python tools/trace_fused.py [bf16x3 tf32x3 ...]
"""
import sys, torch
sys.path.insert(0, '/root/repo')
from sherf_b200 import synthetic as S, _lib
from sherf_b200.triplane import hot_path_modules
dev = torch.device('cuda:0')
model = S.make_smpl_model(0)
sc = S.make_scene(S.SceneSpec(H=512, W=512, samples=64, seed=0), model)
def mv(x):
    if torch.is_tensor(x): return x.to(dev)
    if isinstance(x, dict): return {k: mv(v) for k, v in x.items()}
    if isinstance(x, list): return [mv(v) for v in x]
    return x
sc = {k: mv(v) for k, v in sc.items()}
for prec in (sys.argv[1:] or ['bf16x3', 'tf32x3']):
    ren, dec = hot_path_modules(model, seed=0, mlp_precision=prec, dense_sigma=True)
    ren, dec = ren.to(dev), dec.to(dev)
    lib = _lib.load()
    tr = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    lib.sherf_debug_set_trace(tr.data_ptr())
    for _ in range(2):
        ren(sc['planes'], sc['obs_input_img'], sc['obs_input_feature'], sc['volumes'], None, sc['obs_sp_input'], dec, sc['ray_origins'],
            sc['ray_directions'], sc['near'], sc['far'], sc['input_data'], sc['rendering_options'])
    torch.cuda.synchronize()
    t = tr.view(148, 16).double().mean(0)
    names = ['prod_wait_empty', 'prod_total', 'iss_wait_operand', 'iss_wait_full', 'iss_total', 'epi_wait_acc', 'epi_xload', 'epi_total', 'iss_wait_x', 'iss_mma', 'iss_commit', '-', '-', '-', '-', '-']
    # counters are those of the LAST decoder launch of the view (the 104 879-point tail chunk: 820 tiles)
    print(prec, {n: int(v) for n, v in zip(names, t.tolist())})
    lib.sherf_debug_set_trace(None)
