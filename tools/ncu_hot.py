"""Top stall-sample SASS lines of one kernel in an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_hot.py report.ncu-rep kernel_regex [n_lines] [instance]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--kernel-name', 'regex:' + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# several kernels may follow each other: split on the "Kernel Name" header rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
inst = int(sys.argv[4]) if len(sys.argv) > 4 else 0
b = blocks[inst]
hdr = b['rows'][0]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in b['rows'][1:] if len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']] or 0) for r in body)
texec = sum(int(r[ix['Instructions Executed']] or 0) for r in body)
print(b['name'][:80], 'samples', tot, 'warp-instr', texec)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
top = sorted(body, key=lambda r: -int(r[ix['# Samples']] or 0))[:n]
for r in sorted(top, key=lambda r: int(r[ix['Address']], 16) if r[ix['Address']].startswith('0x') else 0):
    s = int(r[ix['# Samples']] or 0)
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{r[ix['Address']][-5:]} {100*s/tot:5.1f}%  ex={r[ix['Instructions Executed']]:>8}  {r[ix['Source']][:70]:70s} {st}")
