#!/usr/bin/env python
"""Static SASS mnemonic counts per kernel of the sm_100a build (cuobjdump -sass): which kernels issue FP64 tensor MMAs
(DMMA), TMA tensor loads (UTMALDG) with mbarriers (SYNCS), cp.async (LDGSTS), warp shuffles.
Usage: python profiles/sass_evidence.py [path/to/libsfb_b200.so] > profiles/X_sass_evidence.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "sphericalfourierbesseldecompositions.jl_b200/libsfb_b200.so"
WANT = ["DMMA", "UTMALDG", "SYNCS", "LDGSTS", "SHFL", "DFMA", "ATOMS", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    counts, order, cur, it = {}, [], None, iter(names)
    for line in sass.splitlines():
        if "Function :" in line:
            cur = next(it)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for w in WANT:
                if op.startswith(w):
                    counts[cur][w] += 1
    print("# SASS mnemonic counts per kernel of libsfb_b200.so (cuobjdump -sass, sm_100a)")
    print("# DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64); UTMALDG = TMA tensor load (cp.async.bulk.tensor); "
          "SYNCS = mbarrier; LDGSTS = cp.async\n")
    for n in order:
        c = counts[n]
        print(f"{n[:118]:118s} " + " ".join(f"{w}={c[w]:4d}" for w in WANT))


if __name__ == "__main__":
    main()
