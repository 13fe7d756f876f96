"""TEST INFRASTRUCTURE -- recipe that materialises oracle/_ref/ from /root/reference (run by __graft_entry__.build() wherever the
reference tree is mounted; oracle/_ref/ is git-ignored, never committed, and travels to the GPU box with the repo snapshot like the
built .so files).

    python -m oracle.make_ref

The reference path is Python: there is nothing to compile.  What this recipe does instead is copy, byte for byte and into the same
relative layout, exactly the reference files that `import training.triplane` / `training.volumetric_rendering.renderer` load under the
oracle's shims (oracle/ref_shim.py) -- the closure is computed by importing them here and listing the modules whose files live under
/root/reference -- so that on the GPU box, where /root/reference does not exist,
  * `bench.py --impl reference` and the bench's `cpu_baseline` leg time the REFERENCE's own ImportanceRenderer.forward / NeRFDecoder
    on the host cores (`cpu_baseline.kind = "reference"`), and
  * `bench.py --gpu-eager-baseline` times the same unmodified code as eager PyTorch on the B200.
Nothing under sherf_b200/ may import oracle/ (tests/test_abi.py greps for it).
"""
from __future__ import annotations

import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = '/root/reference'
DST_ROOT = os.path.join(HERE, '_ref')


def main() -> int:
    if not os.path.isdir(os.path.join(SRC_ROOT, 'sherf', 'training', 'volumetric_rendering')):
        print('oracle/make_ref: /root/reference not mounted here; keeping whatever oracle/_ref already holds')
        return 0
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_shim
    from sherf_b200 import synthetic as S
    assert ref_shim.REF_ROOT.startswith(SRC_ROOT), 'the recipe must import the mounted reference, not a previous copy'
    ref_shim.load(S.smpl_model_to_torch(S.make_smpl_model(0)))
    files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())
                    if getattr(m, '__file__', None) and os.path.realpath(m.__file__).startswith(SRC_ROOT + os.sep)})
    if os.path.isdir(DST_ROOT):
        shutil.rmtree(DST_ROOT)
    for f in files:
        dst = os.path.join(DST_ROOT, os.path.relpath(f, SRC_ROOT))
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(f, dst)
    with open(os.path.join(DST_ROOT, 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': SRC_ROOT, 'files': [os.path.relpath(f, SRC_ROOT) for f in files],
                   'note': 'byte-identical copies made by oracle/make_ref.py; test infrastructure, git-ignored'}, fh, indent=1)
    print(f'oracle/make_ref: {len(files)} reference files -> {DST_ROOT}')
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
