#!/usr/bin/env python
"""Per-role stall table of the tensor-core triangle kernel from an ncu report captured with --import-source on
(library built with -lineinfo):

  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_src.csv
  python tools/ncu_roles.py X_src.csv > profiles/ncu_triangles_mma_roles_rNN.txt

Sampled warp stalls are summed over the CUDA source lines of each role.  Helper functions are attributed by name
(expand_quad -> expansion; push16 / flush_keys / raise_threshold -> epilogue); the barrier-wait helper and the
common.cuh helpers are shared by all roles and listed separately.
"""
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "sac_cot_b200", "csrc", "kernels_triangles_mma.cu")
HELPERS = {"push16": "epilogue (warps 0-7)", "flush_keys": "epilogue (warps 0-7)", "raise_threshold": "epilogue (warps 0-7)",
           "expand_quad": "expansion (warps 9-14)", "mbar_wait_wd": "barrier waits (helper shared by all roles)"}


def line_roles():
    """role of every line of kernels_triangles_mma.cu"""
    role, cur, func = {}, None, None
    for n, line in enumerate(open(SRC).read().splitlines(), 1):
        m = re.match(r"(?:__device__|__global__).*?\b(\w+)\s*\($", line) or re.match(r"__device__ .*?\b(\w+)\(", line)
        if "triangles_mma_kernel(" in line and "__global__" in line:
            func, cur = "kernel", "setup / teardown"
        elif m and func != "kernel":
            func = m.group(1)
        if func == "kernel":
            if "=== epilogue warps ===" in line:
                cur = "epilogue (warps 0-7)"
            elif "=== MMA issuer" in line:
                cur = "MMA issuer (warp 8)"
            elif "=== expansion warps ===" in line:
                cur = "expansion (warps 9-14)"
            elif "no CTA of the pair may exit" in line:
                cur = "setup / teardown"
            role[n] = cur
        else:
            role[n] = HELPERS.get(func, "small helpers (all roles)")
        if line.startswith("}"):
            func = None
    return role


def main(path):
    roles = line_roles()
    agg = defaultdict(lambda: defaultdict(float))
    cur_file, hdr = None, None
    for r in csv.reader(open(path)):
        if len(r) >= 2 and r[0] in ("File Name", "File Path"):
            cur_file = os.path.basename(r[1])
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or not r or not r[0].strip().isdigit():
            continue
        ln = int(r[0])
        name = roles.get(ln, "other") if cur_file == "kernels_triangles_mma.cu" else f"{cur_file} (helpers shared by all roles)"
        for k, h in enumerate(hdr):
            if k >= len(r):
                break
            if h in ("# Samples", "Instructions Executed") or (h.startswith("stall_") and "Not Issued" not in h):
                try:
                    agg[name][h] += float(r[k].replace(",", ""))
                except ValueError:
                    pass
    total = sum(d["# Samples"] for d in agg.values()) or 1.0
    print(f"# {os.path.basename(path)}: sampled warp stalls of triangles_mma_kernel by role "
          f"(ncu source page, cuda+sass correlation; {total:.0f} samples)")
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
        if d["# Samples"] == 0:
            continue
        print(f"\n== {name}: {d['# Samples']:.0f} samples ({100 * d['# Samples'] / total:.1f} % of all), "
              f"{d['Instructions Executed']:.0f} warp instructions")
        st = sorted(((v, h) for h, v in d.items() if h.startswith("stall_") and v > 0), reverse=True)
        ssum = sum(v for v, _ in st) or 1.0
        for v, h in st[:7]:
            print(f"   {h:24s} {v:10.0f}  {100 * v / ssum:5.1f} % of the role's samples")


if __name__ == "__main__":
    main(sys.argv[1])
