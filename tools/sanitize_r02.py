#!/usr/bin/env python
"""Small self-checking cases of the round-2 paths: second-order mode (both S2 paths), the matching front end
(tensor-core sweep + decision + scan, against the exhaustive path), the in-library sharded call on a one-rank
communicator, a two-member device group on one GPU, a forced key-pool overflow.

Written as a compute-sanitizer driver (`compute-sanitizer --tool memcheck python tools/sanitize_r02.py`), but
compute-sanitizer is closed on this GPU pool (the run is refused: "runs under it have left GPUs needing a reset"), so
NO sanitizer run exists for round 2.  What stands in for it: this script run plainly, and the parity tests, which
compare every new path bit for bit with the CPU oracle on small and ragged sizes."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sac_cot_b200 import _abi, synth  # noqa: E402
from sac_cot_b200.api import Group, Registrar, load_library  # noqa: E402

lib = load_library()
p = synth.make_pair(1400, 0.1, 77)
for path in (0, 1):
    with Registrar(lib=lib, compat_mode=_abi.COMPAT_SECOND_ORDER, so_min_common=12, num_edges=256) as g:
        g.set("triangle_path", path)
        print("second-order path", path, g.register(p.src, p.dst)[2])
rng = np.random.default_rng(1)
f = (rng.random((300, 33)) * 10).astype(np.float32)
gd = (rng.random((700, 33)) * 10).astype(np.float32)
gd[100:150] = gd[99]   # overflowing candidate lists -> scan kernel
f[:3] = gd[[99, 120, 5]]
with Registrar(lib=lib) as g:
    nn, _, _ = g.match(f, np.zeros((300, 3), np.float32), gd, np.zeros((700, 3), np.float32))
    print("match", nn[:4].tolist())
    g.set("match_path", 0)
    nn2, _, _ = g.match(f, np.zeros((300, 3), np.float32), gd, np.zeros((700, 3), np.float32))
    assert (nn == nn2).all()
ident = (ctypes.c_ubyte * _abi.COMM_ID_BYTES)()
assert lib.sac_cot_comm_unique_id(ident) == 0
with Registrar(lib=lib) as g:
    g.comm_init(rank=0, world=1, unique_id=bytes(ident))
    print("sharded world 1", g.register_sharded(p.src, p.dst)[2])
pairs = [synth.make_pair(n, 0.1, 90 + k) for k, n in enumerate((300, 640, 129, 900))]
with Group([0, 0], lib=lib) as grp:
    print("group", grp.register_batch([q.src for q in pairs], [q.dst for q in pairs]).inliers.tolist())
with Registrar(lib=lib, tau_compat=2.0) as g:   # dense graph: the key pool overflows once, the library re-runs
    print("overflow + retry", g.register(p.src[:800], p.dst[:800])[2], "retries", g.get("retries"))
print("done")
