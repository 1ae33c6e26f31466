"""tools/cmp_profiles.py -- per-op comparison of two tools/profile_ops.py logs (e.g. MARS_TC_DEBUG=0 vs 2)."""
import re
import sys


def load(f):
    d = {}
    for l in open(f):
        m = re.match(r"op\s+(\d+) L\s*(\d+) (\S+)\s+impl=(\d) mode=(\d) ic=\s*(\d+) oc=\s*(\d+) o=\s*(\d+)x\s*(\d+) k=(\d) fused=(\d)\s+([\d.]+) ms", l)
        if m:
            d[int(m.group(1))] = (m.group(3), int(m.group(6)), int(m.group(7)), int(m.group(8)), int(m.group(10)), float(m.group(12)), int(m.group(4)))
    return d


a, b = load(sys.argv[1]), load(sys.argv[2])
rows = [(op,) + a[op] + (b[op][5],) for op in a if a[op][0] == "conv_i8_nchw" and a[op][6] == 1 and op in b]
rows.sort(key=lambda r: -r[6])
print("op   ic   oc  res k   A(ms)   B(ms)  tiles/CTA  us/tile(B)")
for r in rows[: int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    op, _, ic, oc, oh, k, ms, impl, ms2 = r
    tiles = ((oh * oh + 127) // 128) * 128
    print("%3d %4d %4d %4d %d  %.3f  %.3f   %6.1f  %6.2f" % (op, ic, oc, oh, k, ms, ms2, tiles / 296, ms2 * 1e3 / (tiles / 296)))
print("sum A %.3f  sum B %.3f" % (sum(r[6] for r in rows), sum(r[8] for r in rows)))
