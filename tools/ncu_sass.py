"""tools/ncu_sass.py -- hot-loop view of one launch of an ncu report: SASS grouped by execution count.
usage: python tools/ncu_sass.py report.ncu-rep launch_index [lines]"""
import collections
import csv
import subprocess
import sys

rep, skip = sys.argv[1], int(sys.argv[2])
nlines = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or not r[ix["Instructions Executed"]].isdigit():
        break  # the SASS view comes first; stop at the next view's header
    data.append(r)
tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
samples = sum(int(r[ix["# Samples"]]) for r in data)
print("warp instructions", tot, "static", len(data), "samples", samples)
cnt, ops, smp = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
for r in data:
    e = int(r[ix["Instructions Executed"]])
    cnt[e] += 1
    smp[e] += int(r[ix["# Samples"]])
    sp = r[ix["Source"]].strip().split()
    op = sp[1] if sp[0].startswith("@") else sp[0]
    ops[e][op.split(".")[0]] += 1
for e, n in sorted(cnt.items(), key=lambda kv: -kv[0] * kv[1])[:8]:
    print("exec %9d x %4d instrs = %5.1f%% of issue, %5.1f%% of samples  %s" % (e, n, 100.0 * e * n / tot, 100.0 * smp[e] / max(samples, 1), dict(ops[e].most_common(16))))
if nlines:
    main = max(cnt.items(), key=lambda kv: kv[0] * kv[1])[0]
    for r in [r for r in data if int(r[ix["Instructions Executed"]]) == main][:nlines]:
        print("%5s  %s" % (r[ix["# Samples"]], r[ix["Source"]].strip()[:100]))
