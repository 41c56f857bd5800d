import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from sac_cot_b200 import synth, _abi
from sac_cot_b200.api import Registrar
for cfg, idxs in (("cfg3_3dlomatch_256x5000", (0,1,2,3)), ("cfg2_3dmatch_256x5000", (0,1))):
    for b in idxs:
        p = synth.make_config_pair(cfg, b)
        with Registrar(device=0) as g:
            g.set("keep_debug", 1); g.set("triangle_path", 1)
            g.params.tau_compat = p.tau_compat; g.params.tau_inlier = p.tau_inlier
            R, t, inl = g.register(p.src, p.dst)
            keys = g.debug(0, _abi.DBG_EDGE_KEYS)
            E = int(g.debug(0, _abi.DBG_NUM_EDGES)[0])
            top = g.debug(0, _abi.DBG_TOP_EDGES)
            tmin = int(top[-1] >> np.uint64(32)) if len(top) else -1
            kmin = int(keys.min() >> np.uint64(32)) if len(keys) else -1
            print(cfg, b, "inliers", inl, "E", E, "kept", len(keys), "min kept T", kmin, "K_e-th T", tmin, flush=True)
