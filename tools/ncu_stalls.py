"""tools/ncu_stalls.py -- top stall sites of one launch of an ncu report (SASS view).
usage: python tools/ncu_stalls.py report.ncu-rep launch_index [n]"""
import csv
import subprocess
import sys

rep, skip = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or not r[ix["Instructions Executed"]].isdigit():
        break
    data.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
tot = sum(agg.values())
print("samples", tot, [(k[6:], round(100.0 * v / tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
    st = sorted(((h[6:], int(r[ix[h]])) for h in stalls if int(r[ix[h]]) > 0), key=lambda kv: -kv[1])[:2]
    print("%6s %9s  %-60s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:60], st))
