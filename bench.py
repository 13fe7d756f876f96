#!/usr/bin/env python
"""Bench of the SHERF render hot path (ImportanceRenderer.forward + NeRFDecoder + ray marcher) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path over N_gpus novel views of BASELINE.json configs[1]
(512x512 RenderPeople-shape, 64 samples/ray, one subject / one observation), synthetic seeded inputs.
With N > 1 (torchrun, one rank per GPU) the step's N x 262 144 rays are sharded across the ranks -- by default at view
granularity (rank r renders view r: the shard of the ray batch it owns), and the step ends with ONE all-gather of the
rendered tiles so that every rank holds all N images (weak scaling: N views on N GPUs).  `--shard tiles` instead
deals every view's rays to all ranks in interleaved 256-ray tiles with one all-gather per view (the single-view
latency mode of sherf_b200.dist.render_sharded).
Prints one JSON line (see README / DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, S = 512, 512, 64
WORKLOAD = 'configs[1]: 512x512 RenderPeople-shape, 64 samples/ray, 1 subject novel view'
FLOP_PER_POINT = 429_248          # MLP MACs x 2 per decoded (surviving) sample
# the fused tcgen05 decoder kernel covers pts_linears[0..7], feature_linear, alpha_linear, views_linear, rgb_linear (triplane.py:293-314)
FLOP_PER_POINT_FUSED = 2 * (71 * 128 + 4 * 128 * 128 + 199 * 128 + 2 * 128 * 128 + 128 * 128 + 128 + 187 * 64 + 64 * 3)
GATHER_BYTES_PER_POINT = 8752     # tri-plane 1536 + 2-D feature 1024 + rgb 48 + 3-D pyramid 6144 bytes of taps per surviving sample
TF32_OVER_BF16 = 0.5              # dense TF32 tensor peak is half the bf16 peak (B200_PROFILING.md table: 1.1 vs 2.25 PF)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='sherf_b200', choices=['sherf_b200', 'reference'])
    ap.add_argument('--precision', default='bf16x3', choices=['fp32', 'tf32', 'tf32x3', 'bf16x3'],
                    help="MLP arithmetic: tf32x3 = error-compensated 3xTF32 on tcgen05 (fp32-grade parity, default); fp32 = CUDA cores")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--gpu-eager-baseline', action='store_true',
                    help='also time oracle/port.py (the reference path written as eager PyTorch) on this GPU for one full view (BASELINE.md 3.4)')
    ap.add_argument('--importance', type=int, default=0,
                    help='fine (importance) samples per ray; 64 = BASELINE configs[4] on one GPU (not the headline workload: the default is configs[1])')
    ap.add_argument('--shard', default='views', choices=['views', 'tiles'], help='N>1: ray-batch sharding granularity')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tensor_tflops': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'src': 'measured (MEASURED_PEAKS.json, bf16 sustained)'}
    return {'hbm_gbs': 6650.0, 'tensor_tflops': 1400.0, 'src': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20',
                                          '-i', str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def make_views(n_views, model):
    """One subject / observation, n_views novel target cameras.  Returns (base scene, list of per-view ray dicts), CPU tensors."""
    from sherf_b200 import synthetic as SY
    base = SY.make_scene(SY.SceneSpec(H=H, W=W, samples=S, seed=0, cam_azim_deg=25.0), model)
    views = [{k: base[k] for k in ('ray_origins', 'ray_directions', 'near', 'far')}]
    for v in range(1, n_views):
        sc = SY.make_scene(SY.SceneSpec(H=H, W=W, samples=S, seed=0, cam_azim_deg=25.0 + 360.0 * v / n_views), model, rays_only=True)
        views.append({k: sc[k] for k in ('ray_origins', 'ray_directions', 'near', 'far')})
    return base, views


def cpu_port_rate(model, n_rays_target, threads, base=None, n_importance=0):
    """Times oracle/port.py (CPU restatement of the reference path) on a strided subset of the same 512x512x64 rays.
    Returns (rate, seconds, description, ray indices, oracle outputs of those rays)."""
    import torch
    from sherf_b200 import synthetic as SY
    from sherf_b200 import dist as sd
    from sherf_b200.triplane import hot_path_modules
    from oracle import port
    torch.set_num_threads(threads)
    if base is None:
        base = SY.make_scene(SY.SceneSpec(H=H, W=W, samples=S, seed=0), model)
    stride = max(1, int(round((H * W / n_rays_target) ** 0.5)))
    idx = (torch.arange(0, H, stride)[:, None] * W + torch.arange(0, W, stride)[None, :]).reshape(-1)
    sub = dict(base)
    for k in ('ray_origins', 'ray_directions', 'near', 'far'):
        sub[k] = base[k][:, idx].contiguous()
    sub['rendering_options'] = dict(base['rendering_options'], depth_resolution_importance=n_importance)
    u = torch.rand(idx.numel(), n_importance, generator=torch.Generator().manual_seed(0)) if n_importance else None
    ren, dec = hot_path_modules(model, seed=0, dense_sigma=True)
    wts = port.hot_path_state_dict(ren, dec)
    mt = SY.smpl_model_to_torch(model)
    clamp = sd.depth_range(base['near'], base['far'], S)             # ray_marcher.py:57 is global over the full view
    t0 = time.perf_counter()
    out = port.render_forward(wts, mt, sub, importance_u=u, depth_clamp=clamp)
    dt = time.perf_counter() - t0
    n = idx.numel() * (S + n_importance)
    desc = f'{idx.numel()} rays (every {stride}th pixel in x and y of the 512x512 view) x {S + n_importance} samples = {n} ray-samples'
    return n / dt, dt, desc, idx, out, u


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference is Python + pytorch3d/spconv (absent) and cannot be
    installed or shipped to the GPU box, so this times oracle/port.py (kind 'port') with all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from sherf_b200 import synthetic as SY
    model = SY.make_smpl_model(0)
    threads = min(os.cpu_count() or 1, 32)      # the chunked brute-force KNN stops scaling (and regresses) beyond ~32 threads
    budget = 150.0 / max(1, args.steps + args.warmup)                 # seconds per step
    n_rays = int(min(16384, max(256, budget * 9000 / S)))             # ~9e3 ray-samples/s/8 cores measured in the build container
    rates, last = [], None
    for i in range(args.warmup + args.steps):
        rate, dt, sample = cpu_port_rate(model, n_rays, threads)[:3]
        if i >= args.warmup:
            rates.append((rate, dt))
        last = sample
    value = sum(r for r, _ in rates) / len(rates)
    line = {
        'impl': 'reference', 'metric': 'ray_samples_per_sec', 'value': value, 'unit': 'ray-samples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(d for _, d in rates) / len(rates), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'H': H, 'W': W, 'samples_per_ray': S, 'step': 'bounded sample: ' + last},
        'cpu_baseline': {'value': value, 'unit': 'ray-samples/s', 'cores': threads, 'kind': 'port', 'sample': last},
        'e2e': {'value': value, 'unit': 'ray-samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from sherf_b200 import synthetic as SY, _lib
    from sherf_b200 import dist as sd
    from sherf_b200.triplane import hot_path_modules

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    lib.sherf_set_profiling(1)
    lib.sherf_last_stage_ms.restype = __import__('ctypes').c_float

    model = SY.make_smpl_model(0)
    base, views = make_views(world, model)
    ren, dec = hot_path_modules(model, seed=0, mlp_precision=args.precision, dense_sigma=True)
    ren, dec = ren.to(dev), dec.to(dev)

    def mv(x):
        if torch.is_tensor(x):
            return x.to(dev)
        if isinstance(x, dict):
            return {k: mv(v) for k, v in x.items()}
        if isinstance(x, list):
            return [mv(v) for v in x]
        return x
    scene = {k: mv(v) for k, v in base.items()}
    scene['rendering_options']['depth_resolution_importance'] = args.importance      # draws: torch.rand on the device inside forward (renderer.py:526)
    N = H * W
    clamp = [sd.depth_range(v['near'], v['far'], S) for v in views]
    # this rank's tiles of every view (device resident for `value`, pinned host copies for `e2e`)
    by_tiles = world > 1 and args.shard == 'tiles'
    my_views = list(range(world)) if (by_tiles or world == 1) else [rank]
    idx = sd.shard_indices(N, rank, world) if by_tiles else torch.arange(N)
    shard_host = [{k: views[v][k][:, idx].contiguous().pin_memory() for k in views[v]} for v in my_views]
    shard_dev = [{k: t.to(dev) for k, t in sh.items()} for sh in shard_host]
    pose_host = {k: base['input_data'][k] for k in ('vertices',)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                   # > 126 MB L2

    def render(sh, v):
        return ren(scene['planes'], scene['obs_input_img'], scene['obs_input_feature'], scene['volumes'], None, scene['obs_sp_input'],
                   dec, sh['ray_origins'], sh['ray_directions'], sh['near'], sh['far'], scene['input_data'], scene['rendering_options'],
                   depth_clamp=clamp[v] if by_tiles else None)

    stage_ms = [0.0] * 8
    host_us = [0.0] * 5             # C-side issue / sync / issue / total, and the Python wrapper around it
    launches = [0]
    points = [0]

    def gather_views(local):
        """view-granular sharding: one all-gather of the rendered [N,5] tiles -> every rank holds all `world` images"""
        full = local.new_empty(world * N, 5)
        dist.all_gather_into_tensor(full, local.contiguous())
        return full

    def step_device():
        outs = []
        for j, v in enumerate(my_views):
            t_py = time.perf_counter()
            rgb, depth, acc = render(shard_dev[j], v)
            host_us[4] += (time.perf_counter() - t_py) * 1e6
            for s_ in range(4):
                host_us[s_] += lib.sherf_last_host_us(s_)
            for s_ in range(8):
                stage_ms[s_] += lib.sherf_last_stage_ms(s_)
            launches[0] += ren.last_launches
            points[0] += ren.last_num_points
            local = torch.cat([rgb[0], depth[0], acc[0]], -1)
            outs.append(sd.all_gather_tiles(local, N) if by_tiles else (gather_views(local) if world > 1 else local))
        return outs

    def step_e2e(host_out):
        for j, v in enumerate(my_views):
            sh = {k: t.to(dev, non_blocking=True) for k, t in shard_host[j].items()}
            scene['input_data']['vertices'] = pose_host['vertices'].to(dev, non_blocking=True)
            rgb, depth, acc = render(sh, v)
            local = torch.cat([rgb[0], depth[0], acc[0]], -1)
            if by_tiles:
                host_out[v].copy_(sd.all_gather_tiles(local, N), non_blocking=True)
            elif world > 1:
                full = gather_views(local)
                for vv in range(world):
                    host_out[vv].copy_(full[vv * N:(vv + 1) * N], non_blocking=True)
            else:
                host_out[v].copy_(local, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        """k steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps (outside the events)."""
        tot = 0.0
        for _ in range(k):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / k

    for _ in range(args.warmup):
        step_device()
    # the timed steps run with the library's per-stage event timers OFF (they cost ~60 host-side event records per forward);
    # the stage split reported next to `value` comes from the same number of extra, untimed steps with the timers on
    lib.sherf_set_profiling(0)
    step_device()
    launches[0] = 0
    points[0] = 0
    host_us[:] = [0.0] * 5
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    ms = timed(step_device, args.steps)
    barrier()
    clk = clocks.stop()
    n_launch_timed, n_points_timed = launches[0], points[0]
    host_timed = [h / (args.steps * len(my_views)) for h in host_us]
    lib.sherf_set_profiling(1)
    step_device()
    stage_ms[:] = [0.0] * 8
    for _ in range(args.steps):
        flush.zero_()
        step_device()
    launches[0], points[0] = n_launch_timed, n_points_timed
    lib.sherf_set_profiling(0)
    host_out = [torch.empty(N, 5).pin_memory() for _ in range(world)]
    for _ in range(args.warmup):
        step_e2e(host_out)
    barrier()
    ms_e2e = timed(lambda: step_e2e(host_out), args.steps)
    barrier()
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        pk = peaks()
        samples_per_step = world * N * (S + args.importance)
        calls = args.steps * len(my_views)
        mlp_ms = stage_ms[3] / calls
        p_call = points[0] / calls
        fused_ms = stage_ms[5] / calls
        chunk_cap = int(os.environ.get('SHERF_CHUNK_CAP', 0)) or 524288       # api.cu chunk_cap_limit()
        n_launch = max(1, -(-int(p_call) // chunk_cap))
        issued_per_useful = 1
        if fused_ms > 0 and args.precision == 'bf16x3':
            # dominant kernel: the ping-pong decoder; its MMAs are kind::f16 (bf16), so the peak is the bf16 one
            roof_kernel = 'k_decoder_pp (tcgen05 kind::f16 bf16 split products, whole NeRFDecoder, two 128-point tiles in flight per SM; %d launches per view)' % n_launch
            ach_tflops = p_call * FLOP_PER_POINT_FUSED / (fused_ms * 1e-3) / 1e12
            algo = f'{FLOP_PER_POINT_FUSED} FLOP per surviving sample x {p_call:.0f} samples per view, avg launch {1e3 * fused_ms / n_launch:.0f} us'
            traffic = 312.3e6 * (p_call / n_launch) / 524288.0   # profiles/r1_p_ncu_full_k_decoder_pp.csv: dram 302.9 MB read + 9.3 MB written per 524 288-point launch
            peak, issued_per_useful = pk['tensor_tflops'], 3
            peak_src = pk['src'] + ': dense bf16; useful FLOPs counted once although bf16x3 issues 3 MMAs per product (issued_frac counts all three)'
        elif fused_ms > 0:       # 3xTF32 / TF32 path: the dominant kernel is the fused decoder trunk
            roof_kernel = 'k_decoder_fused (tcgen05 kind::tf32, whole NeRFDecoder: pts_linears 0-7, feature/alpha, views, rgb; %d launches per view)' % n_launch
            ach_tflops = p_call * FLOP_PER_POINT_FUSED / (fused_ms * 1e-3) / 1e12
            algo = f'{FLOP_PER_POINT_FUSED} FLOP per surviving sample x {p_call:.0f} samples per view, avg launch {1e3 * fused_ms / n_launch:.0f} us'
            traffic = 87.7e6      # profiles/r1_n_ncu_full_k_decoder_fused.csv: dram read 85.3 MB + write 2.4 MB per 131 072-point launch
            peak, issued_per_useful = pk['tensor_tflops'] * TF32_OVER_BF16, (3 if args.precision == 'tf32x3' else 1)
            peak_src = pk['src'] + ' x 0.5: dense TF32 rate is half the bf16 rate; useful FLOPs counted once although tf32x3 issues 3 MMAs per product'
        else:
            roof_kernel = 'MLP stage (k_sgemm fp32 CUDA-core layers)'
            ach_tflops = p_call * FLOP_PER_POINT / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
            algo = f'{FLOP_PER_POINT} FLOP per surviving sample x {p_call:.0f} samples per view'
            traffic = None
            peak = pk['tensor_tflops'] * TF32_OVER_BF16
            peak_src = pk['src'] + ' x 0.5 (TF32 tensor peak, for reference: this path runs on the CUDA cores)'
        h2d = (sum(t_.numel() * 4 for sh in shard_host for t_ in sh.values()) + pose_host['vertices'].numel() * 4 * len(my_views)) * (1 if by_tiles or world == 1 else world)
        line = {
            'metric': 'ray_samples_per_sec', 'value': samples_per_step / (ms * 1e-3), 'unit': 'ray-samples/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': {'fp32': 'f32', 'tf32': 'tf32', 'tf32x3': 'tf32x3 (3xTF32 split products, fp32 accumulate; fp32-grade)',
                                                'bf16x3': 'bf16x3 decoder (bf16 hi/lo split products, 16 significand bits, fp32 accumulate) + tf32x3 fusion/transformer'}[args.precision],
            'data': 'synthetic',
            'config': {'workload': WORKLOAD if not args.importance else f'configs[4] shape on this GPU count: 512x512, {S}+{args.importance} coarse+fine importance sampling',
                       'H': H, 'W': W, 'samples_per_ray': S, 'importance_samples_per_ray': args.importance, 'views_per_step': world,
                       'parallelism': ('single GPU' if world == 1 else (f'256-ray tiles of every view dealt to {world} ranks, one all-gather per view' if by_tiles else
                                        f'ray batch sharded at view granularity over {world} ranks (1 view each), one all-gather of the rendered tiles per step')),
                       'mlp_precision': args.precision, 'surviving_points_per_view': p_call * world if by_tiles else p_call,
                       'l2': 'explicit 256 MiB flush between timed steps (outside the events); working set > 126 MB L2'},
            'e2e': {'value': samples_per_step / (ms_e2e * 1e-3), 'unit': 'ray-samples/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': world * N * 5 * 4, 'ms_per_step': ms_e2e,
                    'note': 'per step: pinned-host rays/near/far/vertices -> device, ImportanceRenderer.forward via the C ABI, rendered rgb+depth+acc -> pinned host'},
            'gpu_launches': launches[0],
            'host_us_per_view_call': {'c_issue_until_sync': host_timed[0], 'c_blocked_in_sync': host_timed[1], 'c_issue_point_stages': host_timed[2],
                                      'c_total': host_timed[3], 'python_forward_total': host_timed[4]},
            'clocks': clk,
            'stages_ms_per_view_call': {n: stage_ms[i] / calls for i, n in enumerate(['prologue+layout', 'cull+compact', 'warp+gather', 'mlp', 'composite', 'mlp:fused_decoder_kernel', 'mlp:fused_transformer_kernel', 'mlp:fused_fusion_kernel'])},
            'mlp_stage_tflops': p_call * FLOP_PER_POINT / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0,
            'roofline': {'bound': 'tensor', 'kernel': roof_kernel, 'achieved': ach_tflops, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': ach_tflops / peak, 'issued_frac': issued_per_useful * ach_tflops / peak, 'traffic': traffic,
                         'peak_source': peak_src, 'algorithmic': algo},
            'decoded_samples_per_sec': points[0] / args.steps * world / (ms * 1e-3),      # surviving samples through the MLP stack (rank 0's count x ranks)
            'gather_hbm': {'achieved_GBps': p_call * GATHER_BYTES_PER_POINT / (stage_ms[2] / calls * 1e-3) / 1e9 if stage_ms[2] > 0 else None,
                           'peak_GBps': pk['hbm_gbs'], 'note': 'SURVEY 8(d) algorithmic 8752 B of feature taps per surviving sample / warp+gather stage time (taps are mostly L2 hits, so this can exceed what DRAM moves)'},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = min(os.cpu_count() or 1, 32)
            rate, dt, sample, idx, want, u_sub = cpu_port_rate(model, 4096, threads, base, args.importance)
            line['cpu_baseline'] = {'value': rate, 'unit': 'ray-samples/s', 'cores': threads, 'kind': 'port', 'sample': sample, 'seconds': dt}
            # BASELINE.json's second metric: PSNR of our image against the oracle's, on the rays the oracle just rendered
            # (outside the timed region; the full 262 144-ray oracle image would take the CPU ~20 minutes)
            import math
            u_full = None
            if args.importance:
                u_full = torch.rand(N, args.importance, device=dev)
                u_full[idx.to(dev)] = u_sub.to(dev)
            got = ren(scene['planes'], scene['obs_input_img'], scene['obs_input_feature'], scene['volumes'], None, scene['obs_sp_input'], dec,
                      shard_dev[0]['ray_origins'], shard_dev[0]['ray_directions'], shard_dev[0]['near'], shard_dev[0]['far'],
                      scene['input_data'], scene['rendering_options'], importance_u=u_full)[0][0].cpu()[idx]
            ref_img = want[0][0]
            hit = base['mask_at_box'].reshape(-1)[idx].bool()

            def psnr(a, b):
                mse = float(((a - b) ** 2).mean()) if a.numel() else 0.0
                return round(10 * math.log10(4.0 / max(mse, 1e-20)), 2)               # images span (-1, 1)
            line['psnr_vs_oracle'] = {'all_pixels_db': psnr(got, ref_img), 'mask_at_box_db': psnr(got[hit], ref_img[hit]),
                                      'pixels': int(idx.numel()), 'mask_at_box_pixels': int(hit.sum()),
                                      'rgb_linf': float((got - ref_img).abs().max()),
                                      'note': 'our render vs oracle/port.py on the cpu_baseline ray sample (test_loop.py:36-37,222-223 metric)'}
        if world == 1 and args.gpu_eager_baseline:
            # "the reference on the same box in GPU-eager mode".  The reference itself cannot travel (pytorch3d, spconv,
            # /root/reference absent on the GPU box), so this is its restatement oracle/port.py run on CUDA tensors: brute-force KNN in
            # 8192-query chunks, torch.inverse per point, F.grid_sample gathers, ~25 small GEMMs -- reported, not optimised against.
            from oracle import port
            wts = {k: v.to(dev) for k, v in port.hot_path_state_dict(ren, dec).items()}
            mt = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in SY.smpl_model_to_torch(model).items()}
            sc_full = dict(scene)
            for k in ('ray_origins', 'ray_directions', 'near', 'far'):
                sc_full[k] = shard_dev[0][k]
            sc_full['rendering_options'] = dict(scene['rendering_options'], depth_resolution_importance=0)
            port.render_forward(wts, mt, sc_full)                                   # warm-up (cuBLAS / cuSOLVER handles, allocator)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ref_out = port.render_forward(wts, mt, sc_full)
            e1.record()
            e1.synchronize()
            secs = e0.elapsed_time(e1) * 1e-3
            ours = ren(scene['planes'], scene['obs_input_img'], scene['obs_input_feature'], scene['volumes'], None, scene['obs_sp_input'], dec,
                       shard_dev[0]['ray_origins'], shard_dev[0]['ray_directions'], shard_dev[0]['near'], shard_dev[0]['far'],
                       scene['input_data'], sc_full['rendering_options'])
            line['gpu_eager_baseline'] = {'value': N * S / secs, 'unit': 'ray-samples/s', 'seconds_per_view': secs, 'kind': 'port',
                                          'what': 'oracle/port.py (restatement of the reference path) as eager PyTorch on this GPU, one full 512x512x64 view',
                                          'peak_mem_GB': torch.cuda.max_memory_allocated() / 1e9,
                                          'rgb_linf_vs_cuda_path_full_view': float((ours[0] - ref_out[0]).abs().max()),
                                          'acc_linf_vs_cuda_path_full_view': float((ours[2] - ref_out[2]).abs().max()),
                                          'rays_beyond_1e-4': float(((ours[0] - ref_out[0]).abs().amax(-1) > 1e-4).float().mean())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
