#!/usr/bin/env python
"""One call of the matching front end on a small batch (for ncu / sanitizer runs): tools/match_one.py [pairs] [N]."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sac_cot_b200 import synth  # noqa: E402
from sac_cot_b200.api import Registrar  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
ps = [synth.make_pair(N, 0.05, 100 + b) for b in range(pairs)]
ds = [synth.make_descriptors(p, 33, seed=b) for b, p in enumerate(ps)]
with Registrar() as reg:
    args = ([d[0] for d in ds], [p.src for p in ps], [d[1] for d in ds], [p.dst for p in ps])
    reg.match_batch(*args)   # warm-up: the first launch of a kernel pays its lazy module load
    reg.set("stage_timing", 1)
    if os.environ.get("MATCH_DBG"):
        reg.set("match_dbg", 1)
    for _ in range(3):
        nn, cs, cd, offs = reg.match_batch([d[0] for d in ds], [p.src for p in ps], [d[1] for d in ds], [p.dst for p in ps])
    print({s: reg.get(f"stage_us_{s}") / 3 for s in ("match_prep", "match_sweep", "match_exact")})
    print("inliers matched:", np.mean([(nn[offs[b]:offs[b + 1]][p.inlier_idx] == p.inlier_idx).mean() for b, p in enumerate(ps)]))
