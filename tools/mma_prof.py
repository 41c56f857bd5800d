#!/usr/bin/env python
"""Experiments: one batch through the tensor-core triangle kernel with the in-kernel wait-time printout."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sac_cot_b200 import synth  # noqa: E402
from sac_cot_b200.api import Registrar  # noqa: E402

dbg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 64
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ps = [synth.make_config_pair("cfg2_3dmatch_256x5000", b) for b in range(pairs)]
with Registrar(device=0) as reg:
    reg.set("triangle_path", 1)
    reg.set("lanes", 1)
    reg.set("tile_runs", runs)
    reg.set("chunk_pairs", pairs)
    reg.register_batch([p.src for p in ps], [p.dst for p in ps])  # warm-up (workspace growth)
    reg.set("triangle_dbg", dbg)
    res = reg.register_batch([p.src for p in ps], [p.dst for p in ps])
    print("inliers[:4]", res.inliers[:4].tolist(), flush=True)
