'''aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line
usage: ncu_lines.py dump.csv [top]'''
import csv, sys, collections
fn = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(fn)))
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ''])   # (file,line) -> inst, samples, thread inst, src
cur_line = None
for r in rows:
    if not r: continue
    if r[0] in ('File Name', 'File Path'): cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = {n: i for i, n in enumerate(r)}; ia = r.index('Address'); continue
    if r[0] in ('Kernel Name', 'Function Name') or hdr is None: continue
    if r[0] != '':
        cur_line = (cur_file, int(r[0])); agg[cur_line][3] = r[1].strip()[:110]
    try:
        inst = int(r[hdr['Instructions Executed']] or 0); smp = int(r[hdr['# Samples']] or 0)
        ti = int(r[hdr['Thread Instructions Executed']] or 0)
    except (ValueError, IndexError):
        continue
    if r[ia] != '' and cur_line:
        a = agg[cur_line]; a[0] += inst; a[1] += smp; a[2] += ti
tot_i = sum(a[0] for a in agg.values()) or 1; tot_s = sum(a[1] for a in agg.values()) or 1
print(f'total inst {tot_i:.3e} samples {tot_s}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{k[0]}:{k[1]:4d} inst {100*a[0]/tot_i:5.1f}%  smp {100*a[1]/tot_s:5.1f}%  lanes {a[2]/max(a[0],1):4.1f}  {a[3]}')
