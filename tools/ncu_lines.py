"""Top source lines by warp-stall samples (or by another per-line column, e.g. "L1 Wavefronts Shared Excessive") for one
kernel of an .ncu-rep captured with --import-source on (development aid).
    python tools/ncu_lines.py gpurun_out/x.ncu-rep factor_kernel [top] [column]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
column = sys.argv[4] if len(sys.argv) > 4 else '# Samples'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name',
                      'regex:' + kern], capture_output=True, text=True).stdout
fpath, hdr, rows = None, None, []


def num(v):
    try:
        return int(float(v))
    except (TypeError, ValueError):
        return 0


for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == 'File Path':
        fpath = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and r[0] and r[0].isdigit() and len(r) > 7:
        d = dict(zip(hdr[4:], r[4:]))
        rows.append((num(d.get(column)), num(d.get('Instructions Executed')), fpath, int(r[0]), r[1].strip()[:110], d))
tot = sum(x[0] for x in rows) or 1
print(f'{kern}: {tot} total of "{column}"')
rows.sort(key=lambda x: -x[0])
for s, ie, f, ln, src, d in rows[:top]:
    st = sorted(((num(v), k) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k), reverse=True)[:3]
    print(f'{100 * s / tot:5.1f}%  {f}:{ln:<4d} {src}   [{", ".join(f"{k[6:]} {v}" for v, k in st if v)}]')
