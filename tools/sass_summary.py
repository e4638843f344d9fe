#!/usr/bin/env python3
"""Summarise the SASS of one kernel: run-lengths of 128-bit loads vs. arithmetic between them.

    python tools/sass_summary.py <cubin-or-so> <substring of the mangled kernel name>

Used to check (before spending GPU time) that ptxas keeps the gather loads batched -- the
number of LDG.E.128 issued back-to-back is the memory-level parallelism one lane can reach.
"""
import re
import subprocess
import sys


def main():
    path, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", txt)
    for blk in blocks[1:]:
        name = blk.split("\n", 1)[0].strip()
        if pat not in name:
            continue
        ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk)
        runs = []
        for op in ops:
            if op.startswith("LDG") and ".128" in op:
                k = "LDG128"
            elif op.startswith("LDG"):
                k = "ldg"
            elif op.startswith(("RED", "ATOM")):
                k = "RED"
            elif op.startswith("SHFL"):
                k = "SHFL"
            elif op.startswith(("STG",)):
                k = "STG"
            elif op.startswith("BRA"):
                k = "BRA"
            else:
                k = "."
            if runs and runs[-1][0] == k:
                runs[-1][1] += 1
            else:
                runs.append([k, 1])
        print(name)
        print("  total instr:", len(ops))
        print("  " + " ".join(f"{k}x{n}" if k != "." else f".{n}" for k, n in runs))


if __name__ == "__main__":
    main()
