#!/usr/bin/env python3
"""Writes profiles/r02/sass_excerpts.txt: per kernel of interest, the count of the SASS mnemonics that prove how it is fed
(bulk copies, mbarriers, cp.async, FP64 tensor-core MMA) and the first few of each, from `cuobjdump -sass` of the built library.
    python profiles/make_sass_excerpts.py"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mpc_fatigue_b200", "lib", "libmpcf.so")
KERNELS = [
    ("k_chain_rule_tmaILi6ELi6ELi5ELi1ELb1", ["UBLKCP", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "LDS.64", "STG.E.EF.64"]),
    ("k_stage_derivsILi6ELi6ELb1", ["LDGSTS", "LDGDEPBAR"]),
    ("k_tree_chain_tcILi40ELi2", ["DMMA", "LDGSTS", "LDS.64", "STS.128", "ATOMG", "STG.E.EF.64", "BAR.SYNC"]),
    ("k_tree_chainILi40", ["LDGSTS", "LDS.128", "REDG", "DFMA"]),
    ("k_forest_fill_crossILi6E7double2", ["STG.E.EF.128"]),
]
HEAD = """SASS excerpts of the sm_100a kernels in mpc_fatigue_b200/lib/libmpcf.so (cuobjdump -sass, CUDA 12.9), round 2; written by
profiles/make_sass_excerpts.py.  What they show: the chain-rule kernel of the static pipeline is fed by the bulk-copy engine
(UBLKCP = cp.async.bulk global->shared) and synchronised with mbarriers (SYNCS.*); the derivative kernel stages M^-1 with LDGSTS
(cp.async); the tree chain-rule kernel issues DMMA.8x8x4 (FP64 mma.sync on the tensor cores), expands the packed stage data with
LDGSTS scatter copies and takes its units from an atomic counter (ATOMG); its DFMA variant reads the matrices with 16-byte
broadcast LDS.128.  No HMMA / UTC*MMA / UTMALDG anywhere: nothing on this path is a low-precision contraction or a tiled tensor
copy.
"""


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    out = [HEAD]
    for key, ops in KERNELS:
        f = next((x for x in funcs if x.split("\n", 1)[0].find(key) >= 0), None)
        if f is None:
            out.append("Function matching %s: not found\n" % key)
            continue
        name, body = f.split("\n", 1)
        ins = [ln for ln in body.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln)]
        out.append("Function : %s" % name.strip())
        out.append("   %d SASS instructions; %s" % (len(ins), ", ".join("%s x%d" % (op, sum(op in ln for ln in ins)) for op in ops)))
        for op in ops:
            for ln in [ln for ln in ins if op in ln][:3]:
                out.append("    " + ln.split(";")[0].rstrip()[8:] + " ;")
        out.append("")
    with open(os.path.join(ROOT, "profiles", "r02", "sass_excerpts.txt"), "w") as fh:
        fh.write("\n".join(out))
    print("\n".join(out[1:12]))


if __name__ == "__main__":
    main()
