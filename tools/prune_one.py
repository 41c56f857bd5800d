#!/usr/bin/env python
"""Diagnostics: one pair through the library with the node-pruning knobs given on the command line.

  SAC_COT_TRACE=1 python tools/prune_one.py N ratio Ke m apex_path node_prune [kitti]

prints the library status, the pruning outcome and the stage times; with SAC_COT_TRACE=1 the library names the
launcher whose kernels failed."""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from sac_cot_b200 import synth  # noqa: E402
from sac_cot_b200.api import Registrar, load_library  # noqa: E402

N, ratio, Ke, m, apex_path, node_prune = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
kw = dict(box=(60.0, 60.0, 6.0), tau_compat=0.6) if len(sys.argv) > 7 and sys.argv[7] == "k" else {}
rect = int(sys.argv[8]) if len(sys.argv) > 8 else 1
reps = 3
p = synth.make_pair(N, ratio, 9300 + N + Ke, **kw)
with Registrar(lib=load_library(), device=0) as reg:
    reg.set("triangle_path", 1)
    reg.set("node_prune", node_prune)
    reg.set("apex_path", apex_path)
    reg.set("node_prune_rect", 2 * rect)
    reg.params.tau_compat = p.tau_compat
    reg.params.tau_inlier = p.tau_inlier
    reg.params.num_edges = Ke
    reg.params.apex_per_edge = m
    reg.register(p.src, p.dst)
    reg.set("stage_timing", 1)
    for _ in range(reps):
        R, t, inl = reg.register(p.src, p.dst)
    names = ("graph", "theta", "triangles", "triangles_kept", "select", "apex", "score")
    print("inliers", inl, "pruned_pairs", reg.get("pruned_pairs"), "kept_nodes", reg.get("kept_nodes"), "rect_pairs", reg.get("rect_pairs"),
          {s: reg.get(f"stage_us_{s}") // reps for s in names}, flush=True)
