"""Per-kernel SASS mnemonic counts of libmm_b200.so (evidence for TMA / mbarrier / packed-FP32 use).

    python tools/sass_summary.py > profiles/rNN_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "audio-mastering-web_b200", "mm_b200", "libmm_b200.so")
KEYS = ["UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "FFMA2", "DFMA", "FFMA", "STG.E.128", "LDS", "STS", "SHFL", "MUFU", "BAR.SYNC"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        mm = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if mm:
            op = mm.group(1)
            counts[fn]["_n"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[fn][k] += 1
    names = list(counts)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    out = ["# SASS mnemonic counts per kernel (`cuobjdump -sass libmm_b200.so`, sm_100a)", "",
           "`UTMALDG` / `UTMASTG` = `cp.async.bulk.tensor` loads / stores (TMA), `SYNCS` = mbarrier operations, `LDGSTS` = per-lane "
           "`cp.async` (edge tiles, aux streams, followers), `FFMA2` = packed FP32.  Template arguments of `sweep2_kernel`: "
           "`<M, NF, NIN, DIR, EPI, NAUX, ST, NF32, NSET>`.", "",
           "| kernel | instr | " + " | ".join(KEYS) + " |", "|---|---|" + "|".join(["---"] * len(KEYS)) + "|"]
    tot = collections.Counter()
    for n, d in zip(names, dem):
        c = counts[n]
        if c["_n"] < 50:
            continue
        short = re.sub(r"\(.*$", "", re.sub(r"^void ", "", d)).replace("mm::", "")
        if len(short) > 110:
            short = short[:107] + "..."
        out.append(f"| `{short}` | {c['_n']} | " + " | ".join(str(c[k]) for k in KEYS) + " |")
        for k in KEYS + ["_n"]:
            tot[k] += c[k]
    out.append(f"| **total** | {tot['_n']} | " + " | ".join(str(tot[k]) for k in KEYS) + " |")
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
