"""tools/tc_probe.py -- run one micro conv model on the GPU and compare with the oracle (debug aid)."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from __graft_entry__ import load_package
from oracle import oraclebind as ob

kw = json.loads(sys.argv[1])
pkg = load_package()
blob = pkg.marsfile.build_single_layer("conv", **kw).to_bytes()
gm = pkg.MarsModel(blob)
om = ob.OracleModel(blob)
rng = np.random.default_rng(2)
W = om.weights_size
fill = rng.integers(0, 256, size=om.arena_bytes - W, dtype=np.uint8)
om.arena()[W:] = fill
gm.mirror()[W:W + fill.size] = fill
gm.arena_upload()
om.run()
print(gm.describe().strip())
rc = gm.run_layer(0)
if rc != 0:
    print("FAIL rc", rc, pkg.lib().mars_b200_last_error())
    sys.exit(1)
got = gm.arena_download()
want = om.arena()[: got.size]
d = np.nonzero(got != want)[0]
print("OK" if d.size == 0 else "MISMATCH %d bytes first %d" % (d.size, d[0]), kw)
