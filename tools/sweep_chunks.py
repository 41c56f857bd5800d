#!/usr/bin/env python
"""Experiments: ms per 256-pair step of the headline workload for a few (lanes, chunk_pairs) settings, device-resident,
CUDA events, one process.  tools/sweep_chunks.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sac_cot_b200 import _abi, synth  # noqa: E402
from sac_cot_b200.api import Registrar  # noqa: E402

import torch  # noqa: E402

W = sys.argv[1] if len(sys.argv) > 1 else "cfg2_3dmatch_256x5000"
cfg = synth.CONFIGS[W]
pairs = cfg["pairs"]
dev = torch.device("cuda", 0)
ps = [synth.make_config_pair(W, b) for b in range(pairs)]
src = np.ascontiguousarray(np.concatenate([p.src for p in ps]))
dst = np.ascontiguousarray(np.concatenate([p.dst for p in ps]))
offsets = np.arange(pairs + 1, dtype=np.int64) * cfg["N"]
d_src, d_dst = torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev)
d_R = torch.empty((pairs, 3, 3), dtype=torch.float32, device=dev)
d_t = torch.empty((pairs, 3), dtype=torch.float32, device=dev)
d_i = torch.empty(pairs, dtype=torch.int32, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
with Registrar(device=0, stream=stream.cuda_stream, tau_compat=cfg["tau"], tau_inlier=cfg["tau"]) as reg:
    step = lambda: reg.register_packed_ptr(d_src.data_ptr(), d_dst.data_ptr(), offsets, d_R.data_ptr(),  # noqa: E731
                                           d_t.data_ptr(), d_i.data_ptr(), _abi.LOC_DEVICE)
    for lanes, chunk in ((3, 0), (2, 128), (3, 128), (4, 64), (3, 64), (2, 86), (4, 0), (3, 0)):
        reg.set("lanes", lanes)
        reg.set("chunk_pairs", chunk)
        for _ in range(3):
            step()
            torch.cuda.synchronize(dev)
            reg.get("last_status")
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for k in range(10):
            flush.fill_(k)
            ev[k][0].record(stream)
            step()
            ev[k][1].record(stream)
        torch.cuda.synchronize(dev)
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        print(f"{W} lanes {lanes} chunk_pairs {chunk or 'auto'}: median {ms[5]:.3f} ms  min {ms[0]:.3f}  "
              f"({pairs / ms[5] * 1e3:.0f} reg/s)", flush=True)
