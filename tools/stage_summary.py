#!/usr/bin/env python
"""Prints value / e2e / per-stage microseconds of bench.py JSON lines (one file per argument)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        j = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    st = j["roofline"]["stage_us_per_step"]
    print(f"{path}: value {j['value']:.0f} e2e {j['e2e']['value']:.0f} ms/step {j['ms_per_step']:.2f} "
          f"clk {j['clocks'].get('sm_mhz')} W {j['clocks'].get('power_w_max')} recall {j['recall_vs_ground_truth']}")
    print("   " + " ".join(f"{k}={v:.0f}" for k, v in st.items()))
