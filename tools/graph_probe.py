#!/usr/bin/env python
"""Tensor-core graph kernel probe: distance error and literal-path rate on one pair of a config."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sac_cot_b200 import synth  # noqa: E402
from sac_cot_b200.api import Registrar  # noqa: E402

for cfg in ("cfg2_3dmatch_256x5000", "cfg4_kitti_128x10000"):
    p = synth.make_config_pair(cfg, 1)
    with Registrar(tau_compat=p.tau_compat, tau_inlier=p.tau_inlier) as g:
        g.set("keep_debug", 1)
        g.set("graph_path", 1)
        g.set("graph_dbg", 1)
        g.register(p.src, p.dst)
        err = g.get("graph_err_e12") * 1e-12
        import math
        print(cfg, "max rel err", err, "= 2^%.1f" % math.log2(err), "groups", g.get("graph_groups"), "unsure", g.get("graph_unsure_groups"),
              "rate", g.get("graph_unsure_groups") / max(1, g.get("graph_groups")))
