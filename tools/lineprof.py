"""Aggregate ncu per-SASS 'Instructions Executed' / stall samples by CUDA source line.
usage: python tools_lineprof.py src_sass.csv cubin kernel_substr [section_index]"""
import csv, re, subprocess, sys, collections
sass_csv, cubin, ksub = sys.argv[1:4]
sec_idx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = list(csv.reader(open(sass_csv)))
secs = []
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        secs.append((i, r[1]))
secs.append((len(rows), None))
want = [k for k in range(len(secs) - 1) if ksub in secs[k][1]]
k = want[sec_idx]
start, end = secs[k][0], secs[k + 1][0]
hdr = rows[start + 1]
body = [dict(zip(hdr, r)) for r in rows[start + 2:end] if len(r) == len(hdr)]
# nvdisasm with line info
out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
# "kernel<(int)8, (int)1>" -> "kernelILi8ELi1EE"; MANGLED=... overrides
mangled = __import__("os").environ.get("MANGLED") or ksub.replace("<(int)", "ILi").replace(", (int)", "ELi").replace(">", "EE")
fn = None; line = None; file = None; seq = []
for l in out:
    if l.startswith(".text."):
        fn = l; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        file, line = m.group(1).split("/")[-1], int(m.group(2)); continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", l) and fn and mangled in fn:
        seq.append((file, line, l.split("*/")[1].strip()))
print("ncu sass rows", len(body), "nvdisasm rows", len(seq), file=sys.stderr)
n = min(len(body), len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for i in range(n):
    f, ln, _ = seq[i]
    ie = int(body[i]["Instructions Executed"] or 0)
    sm = int(body[i]["# Samples"] or 0)
    agg[(f, ln)][0] += ie; agg[(f, ln)][1] += sm; agg[(f, ln)][2] += 1
    tot[0] += ie; tot[1] += sm
src = {}
def getsrc(f, ln):
    import os
    for d in ("/root/repo/bayesformer_b200/csrc/",):
        p = d + f
        if os.path.exists(p):
            if p not in src: src[p] = open(p).read().splitlines()
            return src[p][ln - 1].strip()[:90] if ln - 1 < len(src[p]) else ""
    return ""
print(f"total inst {tot[0]:,} samples {tot[1]:,}")
for (f, ln), (ie, sm, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][int(__import__("os").environ.get("SORTCOL","1"))])[:int(__import__("os").environ.get("TOPN","45"))]:
    print(f"{100*ie/tot[0]:5.1f}% inst {100*sm/max(1,tot[1]):5.1f}% stall  n={cnt:4d}  {f}:{ln}  {getsrc(f, ln)}")
