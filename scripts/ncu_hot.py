"""Developer helper: hottest SASS lines (warp-stall samples) of one kernel in an ncu report."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.012
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# find header row
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for r in rows[h + 1:]:
    if len(r) <= max(i_s, i_ex, i_src) or r[i_src] == "Source":
        break
    try:
        data.append((int(r[i_s] or 0), int(r[i_ex] or 0), r[i_src]))
    except ValueError:
        break
tot = sum(d[0] for d in data) or 1
print("kernel", pat, "total samples", tot, "sass lines", len(data), "inst executed", sum(d[1] for d in data))
for k, (s, ex, src) in enumerate(data):
    if s > tot * thr:
        print(f"{s:6d} {100*s/tot:5.1f}% [{k:4d}] ex={ex:>9d} {src[:110]}")
if len(sys.argv) > 4:
    import collections
    c = collections.Counter()
    for s, ex, src in data:
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
        c[op.split('.')[0]] += ex
    print(c.most_common(25))
