"""Put this directory (and the repo root) on PYTHONPATH and set SHERF_B200_OVERLAY=1: every interpreter started that way --
including the torch.multiprocessing.spawn workers of train.py:98-103 -- resolves `training.triplane` and
`training.volumetric_rendering.renderer` to sherf_b200 (sherf_b200/overlay/__init__.py)."""
import os

if os.environ.get('SHERF_B200_OVERLAY') == '1':
    try:
        from sherf_b200 import overlay as _overlay
        _overlay.install()
    except Exception as _e:                                      # never break interpreter start-up; the import will fail loudly later
        import sys
        sys.stderr.write(f'sherf_b200 overlay not installed: {_e!r}\n')
