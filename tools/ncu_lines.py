"""Per CUDA source line: warp instructions executed and stall samples of one kernel in an ncu report
(needs -lineinfo + --import-source on).  usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [n_lines]"""
import csv, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname, hdr, agg, src = '', None, collections.defaultdict(lambda: [0, 0]), {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path' or r[0] == 'File Name':
        fname = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = {h: i for i, h in enumerate(r)}
        first_src = r.index('Source')
    elif hdr and len(r) > hdr['Instructions Executed'] and r[0].isdigit():  # a CUDA line row carries its SASS rows' totals
        key = (fname, int(r[0]))
        num = lambda x: int(x) if x.isdigit() else 0
        agg[key][0] += num(r[hdr['Instructions Executed']])
        agg[key][1] += num(r[hdr['# Samples']])
        src[key] = r[first_src]
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print('warp-instr', ti, 'samples', ts)
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{key[0][:18]:18s}:{key[1]:4d} instr {100*v[0]/ti:5.1f}%  samples {100*v[1]/max(ts,1):5.1f}%  {src[key].strip()[:100]}")
