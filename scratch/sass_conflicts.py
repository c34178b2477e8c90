'''shared-memory excessive wavefronts per source line (ncu SASS page joined with nvdisasm -g)
usage: python scratch/sass_conflicts.py REPORT KERNEL_REGEX SASS_FILE MANGLED_SUBSTR [N]'''
import csv, re, subprocess, sys
rep, rx, sass, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
amap, infn, cur = {}, False, None
for ln in open(sass):
    if ln.lstrip().startswith('.text.'):
        infn = mangled in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: amap[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
hdr = None; agg = {}; base = None
def I(x):
    try: return int(x)
    except ValueError: return 0
for r in csv.reader(out.splitlines()):
    if r and r[0] == 'Address':
        if hdr is not None: break
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try: addr = int(r[0], 16)
    except ValueError: continue
    if base is None: base = addr
    k, txt = amap.get(addr - base, (None, r[1]))
    a = agg.setdefault(k, [0, 0, 0])
    a[0] += I(r[hdr.index('L1 Wavefronts Shared Excessive')]); a[1] += I(r[hdr.index('L1 Wavefronts Shared')]); a[2] += I(r[hdr.index('L1 Wavefronts Shared Ideal')])
te = sum(a[0] for a in agg.values()); tw = sum(a[1] for a in agg.values())
print('excessive', te, 'of', tw, 'shared wavefronts')
src = {}
def text(k):
    if k is None: return ''
    f, l = k
    if f not in src:
        try: src[f] = open('yanocomp_b200/csrc/' + f).read().splitlines()
        except OSError: src[f] = []
    return src[f][l - 1].strip()[:90] if l <= len(src[f]) else ''
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100 * a[0] / max(te, 1):5.1f}% exc  {a[1] / max(a[2], 1):5.1f}x  {k}  {text(k)}')
