"""`python -m sherf_b200.overlay <script.py> [args...]`: run a reference script (train.py) unchanged with the overlay installed,
in this process and -- through PYTHONPATH + sitecustomize -- in the workers it spawns."""
import os
import runpy
import sys

from . import HERE, install


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    site = os.path.join(HERE, '_site')
    repo = os.path.dirname(os.path.dirname(HERE))
    os.environ['SHERF_B200_OVERLAY'] = '1'
    os.environ['PYTHONPATH'] = os.pathsep.join([site, repo] + [p for p in os.environ.get('PYTHONPATH', '').split(os.pathsep) if p])
    install()
    script = os.path.abspath(sys.argv[1])
    sys.argv = sys.argv[1:]
    sys.path.insert(0, os.path.dirname(script))
    runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
    main()
