#!/usr/bin/env python
"""cuobjdump -sass opcode histogram per kernel of libd2pc_b200.so -> profiles/sass_rN.txt (no GPU needed).

    python tools/sass_histogram.py [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "disparity_to_point_cloud_b200", "libd2pc_b200.so")
WATCH = ("UBLKCP", "UBLKPF", "SYNCS", "REDUX", "VIADDMNMX", "ATOMS", "LDGSTS", "UTMALDG", "HMMA", "STS", "LDS", "STG", "LDG",
         "DFMA", "DMUL", "DADD", "F2F", "MUFU", "BAR", "ATOMG", "RED")


def main(out_path):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op = m.group(1)
            base = op.split(".")[0]
            hist[cur][base] += 1
            if base in WATCH:
                hist[cur]["*" + op] += 1
    tot = collections.Counter()
    with open(out_path, "w") as f:
        f.write("# SASS opcode histograms of libd2pc_b200.so (cuobjdump -sass, sm_100a), one block per kernel.\n"
                "# Lines marked * carry the full mnemonics of the memory / async / FP64 / packed instructions:\n"
                "#   UBLKCP = cp.async.bulk (TMA bulk copy: the band kernel's row loads), UBLKPF = cp.async.bulk.prefetch.L2,\n"
                "#   SYNCS = mbarrier, REDUX = warp reduce (four units per REDUX in the band kernel), VIADDMNMX.S16x2 = packed\n"
                "#   DPX compare (SWAR median), ATOMS = shared-memory atomics (SWAR median variant 3).\n"
                "#   No UTMALDG / UTC*MMA / HMMA: a 1-D streaming path with no contraction uses neither tensor cores nor\n"
                "#   tiled TMA descriptors (north_star).\n")
        for k, h in hist.items():
            base = sorted(((kk, v) for kk, v in h.items() if not kk.startswith("*")), key=lambda x: -x[1])
            star = sorted(((kk[1:], v) for kk, v in h.items() if kk.startswith("*")), key=lambda x: -x[1])
            f.write(f"\n## {k}\n   {sum(v for _, v in base)} instructions\n   " + ", ".join(f"{kk} {v}" for kk, v in base[:24]) + "\n")
            if star:
                f.write("   * " + ", ".join(f"{kk} {v}" for kk, v in star[:30]) + "\n")
            for kk, v in base:
                tot[kk] += v
        f.write("\n## whole library\n   " + ", ".join(f"{kk} {v}" for kk, v in tot.most_common(40)) + "\n")
        for key in ("UBLKCP", "UBLKPF", "SYNCS", "REDUX", "ATOMS", "UTMALDG", "HMMA", "LDGSTS"):
            f.write(f"   {key}: {tot.get(key, 0)}\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_r2.txt"))
