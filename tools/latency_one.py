#!/usr/bin/env python
"""Single-pair latency breakdown: tools/latency_one.py [N] — per-stage device microseconds (CUDA events inside the
library) and the wall clock of the host-buffer call, p50 over 100 calls."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sac_cot_b200 import _abi, synth  # noqa: E402
from sac_cot_b200.api import Registrar  # noqa: E402

import torch  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
p = synth.make_pair(N, 0.05, 42)
hs, hd = torch.from_numpy(p.src).pin_memory(), torch.from_numpy(p.dst).pin_memory()
oR, ot, oi = torch.empty(9).pin_memory(), torch.empty(3).pin_memory(), torch.empty(1, dtype=torch.int32).pin_memory()
off = np.array([0, N], dtype=np.int64)
STAGES = ("pack", "graph", "scan", "theta", "triangles", "select", "apex", "kabsch", "score", "finalize")
with Registrar() as reg:
    call = lambda: reg.register_packed_ptr(hs.data_ptr(), hd.data_ptr(), off, oR.data_ptr(), ot.data_ptr(), oi.data_ptr(), _abi.LOC_HOST)  # noqa: E731
    for _ in range(10):
        call()
    ts = []
    for _ in range(100):
        t0 = time.perf_counter()
        call()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"N={N} wall p50 {1e6 * ts[50]:.1f} us  min {1e6 * ts[0]:.1f} us  inliers {int(oi[0])} launches/call {reg.get('launches') // 110}")
    reg.set("stage_timing", 1)
    for _ in range(50):
        call()
    us = {s: reg.get(f"stage_us_{s}") / 50 for s in STAGES}
    print({k: round(v, 1) for k, v in us.items()}, "sum", round(sum(us.values()), 1))
