"""Summarises `ncu -i X.ncu-rep --page source --csv --print-source sass` output: hot SASS segments and opcode mix.
usage: python scripts/sass_hot.py file.csv [dump_first dump_last]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
it = hdr.index("Thread Instructions Executed")
data = [(r[ia].strip(), int(r[ie]), int(r[isamp]), int(r[it])) for r in rows[hi + 1:] if len(r) > ie and r[ie].isdigit()]
tot = sum(d[1] for d in data); print("total warp instr", tot, "lines", len(data), "avg threads", sum(d[3] for d in data) / tot)
if len(sys.argv) > 3:
    for k in range(int(sys.argv[2]), int(sys.argv[3]) + 1): print(k, data[k][1], data[k][2], data[k][0])
    sys.exit()
segs = []; cur = None
for k, (src, e, s, t) in enumerate(data):
    if cur and abs(e - cur[2]) <= 0.02 * max(e, cur[2]): cur[1] = k; cur[3] += e; cur[4] += s
    else: cur = [k, k, e, e, s]; segs.append(cur)
for s in sorted(sorted(segs, key=lambda s: -s[3])[:25]):
    print(f"lines {s[0]:5d}-{s[1]:5d} n={s[1]-s[0]+1:4d} exec/line={s[2]:9d} share={s[3]/tot:.3f} samples={s[4]}")
ops = collections.Counter()
for src, e, s, t in data:
    tk = src.split(); op = tk[1] if tk[0].startswith('@') else tk[0]; ops[op.split('.')[0]] += e
print("  ".join(f"{op}:{c/tot:.3f}" for op, c in ops.most_common(24)))
