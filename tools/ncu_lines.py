"""Per-source-line hot spots of one kernel in an ncu report captured with --import-source on.

usage: python tools/ncu_lines.py report.ncu-rep kernel-regex [top]
Prints the source lines with the most executed warp instructions and stall samples.
"""
import csv
import io
import subprocess
import sys


def main(path, kernel, top=25):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name', 'regex:' + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    lines, fname, hdr = [], '', None
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        elif len(r) > 8 and r[0] == 'Line No':
            hdr = r
        elif hdr and len(r) > 8 and r[0].isdigit():
            try:
                inst = int(r[hdr.index('Instructions Executed')])
                samp = int(r[hdr.index('# Samples')])
            except ValueError:
                continue
            lines.append((inst, samp, fname, int(r[0]), r[1].strip()[:110]))
    ti, ts = sum(l[0] for l in lines), sum(l[1] for l in lines)
    print(f'{kernel}: {ti} warp instructions, {ts} stall samples over {len(lines)} source lines')
    print('-- by instructions executed')
    for inst, samp, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f'{100 * inst / max(ti, 1):5.1f}% inst {100 * samp / max(ts, 1):5.1f}% samp  {f}:{ln}  {src}')
    print('-- by stall samples')
    for inst, samp, f, ln, src in sorted(lines, key=lambda l: -l[1])[:top]:
        print(f'{100 * inst / max(ti, 1):5.1f}% inst {100 * samp / max(ts, 1):5.1f}% samp  {f}:{ln}  {src}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
