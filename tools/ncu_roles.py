#!/usr/bin/env python
"""Per-role stall table of the tensor-core triangle kernel from an `ncu --page source --csv` export (captured with
--import-source on, library built with -lineinfo): sampled warp stalls summed over the source lines of each role.

  python tools/ncu_roles.py gpurun_out/xyz_src.csv > profiles/ncu_triangles_mma_roles_rNN.txt
"""
import csv
import re
import sys
from collections import defaultdict

ROOT_FILE = "kernels_triangles_mma.cu"


def role_ranges():
    """Line ranges of the three roles, found from the marker comments in the kernel source."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sac_cot_b200", "csrc", ROOT_FILE)
    marks = {}
    for n, line in enumerate(open(path), 1):
        if "=== epilogue warps ===" in line:
            marks["epilogue"] = n
        elif "=== MMA issuer" in line:
            marks["issuer"] = n
        elif "=== expansion warps ===" in line:
            marks["expansion"] = n
        elif "no CTA of the pair may exit" in line:
            marks["end"] = n
    return [("epilogue (warps 0-7)", marks["epilogue"], marks["issuer"]), ("MMA issuer (warp 8)", marks["issuer"], marks["expansion"]),
            ("expansion (warps 9-14)", marks["expansion"], marks["end"])]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    col = {h: k for k, h in enumerate(hdr)}
    src_col = next((col[h] for h in hdr if h.strip().lower() in ("source", "#")), None)
    line_col = next((col[h] for h in hdr if h.strip() in ("#", "Line", "Source Line")), 0)
    stall_cols = [h for h in hdr if h.startswith("stall_") or "Stall" in h or h.startswith("smsp__pcsamp_warps_issue_stalled")]
    samp_col = next((h for h in hdr if h.strip() in ("# Samples", "Sampling Data (All)", "Samples", "Warp Stall Sampling (All Samples)")), None)
    inst_col = next((h for h in hdr if h.strip() in ("Instructions Executed", "inst_executed")), None)
    roles = role_ranges()
    agg = {r[0]: defaultdict(float) for r in roles}
    agg["other (setup, helpers inlined elsewhere)"] = defaultdict(float)
    for r in rows[1:]:
        try:
            ln = int(re.sub(r"[^0-9]", "", r[line_col]) or 0)
        except (ValueError, IndexError):
            continue
        name = next((n for n, a, b in roles if a <= ln < b), "other (setup, helpers inlined elsewhere)")
        for h in stall_cols + ([samp_col] if samp_col else []) + ([inst_col] if inst_col else []):
            try:
                agg[name][h] += float(r[col[h]].replace(",", "") or 0)
            except (ValueError, IndexError):
                pass
    print(f"# {path}: sampled warp stalls by role (source lines of {ROOT_FILE})")
    for name, d in agg.items():
        total = d.get(samp_col, 0.0) if samp_col else sum(d[h] for h in stall_cols)
        print(f"\n== {name}: samples {total:.0f}" + (f", instructions executed {d.get(inst_col, 0):.0f}" if inst_col else ""))
        top = sorted(((v, h) for h, v in d.items() if h in stall_cols and v > 0), reverse=True)[:8]
        for v, h in top:
            print(f"   {h:60s} {v:12.0f}  {100 * v / max(1.0, sum(x for x, _ in top)):5.1f} %")
    if not stall_cols:
        print("(no stall columns found; header was:", hdr[:30], ")")


if __name__ == "__main__":
    main(sys.argv[1])
