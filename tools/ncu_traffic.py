"""profiles/ncu_traffic.json from an `ncu --set full` report: measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per
surviving point of the point-stage kernels, keyed by the library's stage-timer index (bench.py KERNELS).  bench.py multiplies by the points
per launch of ITS run to fill `roofline.traffic` -- measured, not typed in.

usage: python tools/ncu_traffic.py gpurun_out/<tag>_full.ncu-rep <points in the captured launch> <tag> > profiles/ncu_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

STAGE = {'k_front_fused': 2, 'k_decoder_pp': 5, 'k_xformer_bf16': 6}


def main(path, points, tag):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(names)}
    res = {}
    for vals in rows[2:]:
        if len(vals) < len(names):
            continue
        kname = vals[col['Kernel Name']]
        for k, stage in STAGE.items():
            if k in kname and 'k_pack' not in kname and str(stage) not in res:
                def get(metric):
                    v = float(vals[col[metric]].replace(',', ''))
                    u = units[col[metric]].lower()
                    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
                rd, wr = get('dram__bytes_read.sum'), get('dram__bytes_write.sum')
                res[str(stage)] = {'kernel': kname[:80], 'dram_bytes_read': rd, 'dram_bytes_write': wr, 'points_in_launch': points,
                                   'dram_bytes_per_point': (rd + wr) / points,
                                   'duration_us_under_ncu': float(vals[col['gpu__time_duration.sum']].replace(',', '')) / (1e3 if units[col['gpu__time_duration.sum']] == 'ns' else 1),
                                   'source': f'ncu --set full --clock-control none, {tag} ({path.split("/")[-1]}): dram__bytes_read.sum + dram__bytes_write.sum of one {points}-point launch'}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3])
