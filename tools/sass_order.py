"""Where do the Philox draws sit relative to the mbarrier waits in the SASS of mc_forward_kernel<FAST>?
usage: python tools/sass_order.py all.sass [FAST]   (all.sass = nvdisasm --print-line-info of the library's cubin)"""
import re, sys
fast = sys.argv[2] if len(sys.argv) > 2 else "1"
fn = None; cur = None; seq = []
for l in open(sys.argv[1]):
    if l.startswith('.text.'): fn = l.strip(); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if fn and f'mc_forward_kernelILi{fast}' in fn and re.match(r'\s+/\*[0-9a-f]+\*/', l):
        seq.append((cur, l.split('*/')[1].strip()))
print(len(seq), "instructions")
ev = []
for i, (c, ins) in enumerate(seq):
    if 'TRYWAIT' in ins: ev.append((i, 'WAIT'))
    if 'UTCBAR' in ins: ev.append((i, 'COMMIT'))
    if 'BAR.SYNC' in ins: ev.append((i, 'BAR'))
ph = [i for i, (c, ins) in enumerate(seq) if c and c[0] == 'common.cuh' and 60 <= c[1] <= 85 and 'IMAD' in ins]
cl = []
for i in ph:
    if not cl or i - cl[-1][1] > 100: cl.append([i, i, 1])
    else: cl[-1][1] = i; cl[-1][2] += 1
for c in cl: ev.append((c[0], f'PHILOX begin ({c[2]} IMADs, to {c[1]})'))
for e in sorted(ev): print(e)
