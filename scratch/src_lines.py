'''per-source-line share of samples / instructions of one kernel in an ncu report (needs --import-source on)
usage: python scratch/src_lines.py REPORT KERNEL_REGEX [N]'''
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx],
                     capture_output=True, text=True).stdout
cur = hdr = None
agg, tot, seen = {}, 0, 0
def I(x):
    try: return int(x)
    except ValueError: return 0
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1]; continue
    if len(r) == 2 and r[0] == 'Function Name':
        seen += 1; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or seen > 1 and False:
        continue
    try: ln = int(r[0])
    except ValueError: continue
    s, ins = I(r[hdr.index('# Samples')]), I(r[hdr.index('Instructions Executed')])
    k = (cur.split('/')[-1], ln, r[1].strip()[:100])
    a = agg.setdefault(k, [0, 0]); a[0] += s; a[1] += ins; tot += s
ti = sum(a[1] for a in agg.values())
print('samples', tot, 'instructions', ti)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100 * a[0] / tot:5.1f}% smp {100 * a[1] / ti:5.1f}% ins  {k[0]}:{k[1]}  {k[2]}')
