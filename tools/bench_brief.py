"""Print the interesting part of a bench.py JSON line read from stdin (used by the tools/*.sh GPU scripts)."""
import json
import sys

line = sys.stdin.read().strip()
try:
    d = json.loads(line)
    print(' '.join(sys.argv[1:]), 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'],
          {k: round(v, 3) for k, v in d.get('stages_ms_per_view_call', {}).items()}, 'roofline', round(d['roofline']['frac'], 3), 'host_us', {k: round(v) for k, v in d.get('host_us_per_view_call', {}).items()})
except Exception as e:  # noqa: BLE001
    print('unparsable bench output:', line[-400:], e)
