#!/usr/bin/env python
"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck): exercises every kernel,
both triangle kernels (block and chunked rows), ragged batch sizes, two lanes, sharded phases."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sac_cot_b200 import synth  # noqa: E402
from sac_cot_b200.api import Registrar  # noqa: E402


def main():
    big = "--big" in sys.argv
    sizes = (3, 129, 300, 777, 1000) + ((11500,) if big else ())
    pairs = [synth.make_pair(n, 0.2, 100 + k) for k, n in enumerate(sizes)]
    with Registrar(device=0, num_edges=128, apex_per_edge=4) as reg:
        if "--tensor" in sys.argv:  # force S2 onto the tensor-core kernel (default: by edge density)
            reg.set("triangle_path", 1)
        res = reg.register_batch([p.src for p in pairs], [p.dst for p in pairs])
        print("batch inliers", res.inliers.tolist())
        reg.params.score_mode = 1
        res = reg.register_batch([p.src for p in pairs[:3]], [p.dst for p in pairs[:3]])
        print("mode1 inliers", res.inliers.tolist())
        reg.params.score_mode = 0
        p = pairs[3]
        regs = [Registrar(device=0, num_edges=128, apex_per_edge=4) for _ in range(2)]
        ph = [r.sharded_phase1(p.src, p.dst, g, 2) for g, r in enumerate(regs)]
        t_all = np.stack([x[0] for x in ph]); c_all = np.stack([x[1] for x in ph])
        best = max(r.sharded_phase2(t_all, c_all) for r in regs)
        print("sharded", [r.sharded_phase3(best)[2] for r in regs])
        for r in regs:
            r.close()


if __name__ == "__main__":
    main()
