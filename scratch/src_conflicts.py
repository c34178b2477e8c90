'''shared-memory excessive wavefronts per source line from an ncu report WITH imported source
usage: python scratch/src_conflicts.py REPORT KERNEL_REGEX [N]'''
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
cur = hdr = None; agg = {}; nfun = 0
def I(x):
    try: return int(x)
    except ValueError: return 0
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1]; continue
    if len(r) == 2 and r[0] == 'Function Name':
        continue
    if len(r) == 2 and r[0] == 'Kernel Name':
        nfun += 1
        if nfun > 1: break
        continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try: ln = int(r[0])
    except ValueError: continue
    k = (cur.split('/')[-1], ln, r[1].strip()[:90])
    a = agg.setdefault(k, [0, 0, 0, 0])
    a[0] += I(r[hdr.index('L1 Wavefronts Shared Excessive')]); a[1] += I(r[hdr.index('L1 Wavefronts Shared')]); a[2] += I(r[hdr.index('L1 Wavefronts Shared Ideal')]); a[3] += I(r[hdr.index('# Samples')])
te = sum(a[0] for a in agg.values()); tw = sum(a[1] for a in agg.values()); ts = sum(a[3] for a in agg.values())
print('excessive', te, 'of', tw, 'shared wavefronts')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{100 * a[1] / max(tw, 1):5.1f}% wavefronts {100 * a[0] / max(te, 1):5.1f}% exc  {a[1] / max(a[2], 1):4.1f}x  {100 * a[3] / max(ts, 1):4.1f}% smp  {k[0]}:{k[1]}  {k[2]}')
