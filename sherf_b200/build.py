"""Builds libsherf_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libsherf_b200.so')
SOURCES = ['api.cu', 'prologue.cu', 'cull.cu', 'gather.cu', 'mlp_simt.cu', 'mlp_umma.cu', 'decoder_fused.cu', 'decoder_pp.cu', 'xformer_fused.cu', 'xformer_bf16.cu', 'front_fused.cu', 'fusion_fused.cu', 'composite.cu', 'importance.cu', 'rays.cu', 'sparse_encoder.cu', 'observation.cu', 'smpl_forward.cu', 'backward.cu', 'backward_umma.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _newest_source_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'sherf_b200.h')]
    return max(os.path.getmtime(f) for f in files)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    extra = ['-DSHERF_FUSED_TRACE'] if os.environ.get('SHERF_FUSED_TRACE') else []      # cycle-counter tracing build (tools/trace_fused.py)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, *extra, '-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f'--- nvcc {src} ---\n{out}\n')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed building libsherf_b200.so')
    cmd = [nvcc, '-shared', '-o', LIB_PATH, *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
