'''join the SASS page of an ncu report (no imported source) with nvdisasm -g line info of the same cubin
usage: python scratch/sass_lines.py REPORT KERNEL_REGEX SASS_FILE MANGLED_SUBSTR [N]'''
import csv, re, subprocess, sys
rep, rx, sass, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
# 1. address -> (file, line) from nvdisasm
amap, infn, cur = {}, False, None
for ln in open(sass):
    if ln.lstrip().startswith('.text.'):
        infn = mangled in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        amap[int(m.group(1), 16)] = (cur, m.group(2))
print('sass instructions', len(amap))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx],
                     capture_output=True, text=True).stdout
hdr = None; agg = {}; tot = 0; ti = 0; n = 0; base = None; opagg = {}
for r in csv.reader(out.splitlines()):
    if r and r[0] == 'Address':
        if hdr is not None: break          # first kernel instance only
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try: addr = int(r[0], 16)
    except ValueError: continue
    if base is None: base = addr
    off = addr - base
    def I(x):
        try: return int(x)
        except ValueError: return 0
    s, ins = I(r[hdr.index('# Samples')]), I(r[hdr.index('Instructions Executed')])
    k, txt = amap.get(off, (None, r[1]))
    a = agg.setdefault(k, [0, 0]); a[0] += s; a[1] += ins; tot += s; ti += ins; n += 1
    op = r[1].split()[0] if not r[1].startswith('@') else r[1].split()[1]
    o = opagg.setdefault(op.split('.')[0], [0, 0]); o[0] += s; o[1] += ins
print('report instructions', n, 'samples', tot, 'executed', ti)
src = {}
def text(k):
    if k is None: return ''
    f, l = k
    if f not in src:
        try: src[f] = open('yanocomp_b200/csrc/' + f).read().splitlines()
        except OSError: src[f] = []
    return src[f][l - 1].strip()[:100] if l <= len(src[f]) else ''
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100 * a[0] / tot:5.1f}% smp {100 * a[1] / ti:5.1f}% ins  {k}  {text(k)}')
print('--- opcodes')
for k, a in sorted(opagg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f'{100 * a[0] / tot:5.1f}% smp {100 * a[1] / ti:5.1f}% ins  {k}')
# ---- address-range attribution: python scratch/sass_lines.py ... N LO_LINE HI_LINE FILE  (samples between the first and
# last instruction mapped to FILE:[LO,HI], inlined helpers in between included)
if len(sys.argv) > 8:
    lo_l, hi_l, fn = int(sys.argv[6]), int(sys.argv[7]), sys.argv[8]
    addrs = sorted(a for a, (k, _) in amap.items() if k and k[0] == fn and lo_l <= k[1] <= hi_l)
    a0, a1 = addrs[0], addrs[-1]
    hdr = None; s_in = i_in = 0
    for r in csv.reader(out.splitlines()):
        if r and r[0] == 'Address':
            if hdr is not None: break
            hdr = r; continue
        if hdr is None or len(r) < len(hdr): continue
        try: addr = int(r[0], 16) - base
        except ValueError: continue
        if a0 <= addr <= a1:
            s_in += I(r[hdr.index('# Samples')]); i_in += I(r[hdr.index('Instructions Executed')])
    print(f'range {fn}:{lo_l}-{hi_l} = addresses {a0:#x}..{a1:#x}: {100 * s_in / tot:.1f}% of samples, {100 * i_in / ti:.1f}% of instructions')
