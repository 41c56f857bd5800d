#!/usr/bin/env python
"""Opcode evidence from the built product library: `cuobjdump -sass sac_cot_b200/lib/libsaccot.so`, per kernel, the
instruction count and the count of the mnemonics that prove which hardware path the kernel uses —
UTCOMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st: TMEM), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk,
the 1-D TMA form), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (packed fp32x2), POPC, REDUX, LDGSTS (cp.async).

  python tools/sass_summary.py > profiles/sass_summary_rNN.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sac_cot_b200", "lib", "libsaccot.so")
WATCH = ["UTCOMMA", "UTCQMMA", "UTCHMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS",
         "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "POPC", "REDUX", "LDGSTS", "ATOMG", "ATOMS", "RED", "HMMA", "IMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op, mods = m.group(1), m.group(2)
            kernels[cur]["_total"] += 1
            kernels[cur][op] += 1
            if op in ("UTCOMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP"):
                kernels[cur][op + mods] += 1
    names = list(kernels)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, dem):
        demangle[n] = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "")).replace("saccot::", "")
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} — {len(kernels)} kernels, sm_100a")
    print(f"# {'kernel':58s} {'instr':>7s}  watched opcodes (count)")
    for n, c in kernels.items():
        seen = [f"{k}={c[k]}" for k in WATCH if c.get(k)]
        print(f"{demangle[n][:58]:58s} {c['_total']:7d}  {' '.join(seen)}")
        detail = sorted(k for k in c if "." in k)
        if detail:
            print(f"{'':58s} {'':7s}  " + " ".join(f"{k}={c[k]}" for k in detail))
    return 0


if __name__ == "__main__":
    sys.exit(main())
