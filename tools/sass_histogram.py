#!/usr/bin/env python
"""Opcode histogram of one kernel of libai_b200.so (cuobjdump -sass), written to profiles/.

    python tools/sass_histogram.py <mangled-or-substring> <out-name> [--lib path]

Evidence for statements such as "sm_100 packed FP32 (FFMA2)", "128-bit streaming loads", "no TMA":
the histogram is of the SHIPPED binary, per kernel, next to the ncu summary of the same kernel.
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pattern", help="substring of the mangled kernel name, e.g. eval_kernelILi0ELi6EfLi5E")
    ap.add_argument("out")
    ap.add_argument("--lib", default=os.path.join(ROOT, "adaptive_interpolation_b200", "_lib", "libai_b200.so"))
    args = ap.parse_args()
    names = subprocess.run(["cuobjdump", "-elf", args.lib], capture_output=True, text=True).stdout
    funcs = sorted(set(re.findall(r"\.text\.(_Z\w+)", names)))
    match = [f for f in funcs if args.pattern in f]
    if len(match) != 1:
        sys.exit("pattern matches %d kernels: %s" % (len(match), match[:5]))
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", match[0], args.lib], capture_output=True, text=True).stdout
    ops = collections.Counter()
    n = 0
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            ops[m.group(1)] += 1
            n += 1
    demangled = subprocess.run(["cu++filt", match[0]], capture_output=True, text=True).stdout.strip() or match[0]
    base = collections.Counter()
    for op, c in ops.items():
        base[op.split(".")[0]] += c
    lines = ["# SASS opcode histogram: `%s`" % demangled, "",
             "`cuobjdump -sass -fun %s libai_b200.so` (sm_100a), %d instructions." % (match[0], n), "",
             "| opcode (base) | count |", "|---|---|"]
    for op, c in base.most_common():
        lines.append("| %s | %d |" % (op, c))
    lines += ["", "Full mnemonics of the memory and FP instructions:", "", "| mnemonic | count |", "|---|---|"]
    for op, c in sorted(ops.items(), key=lambda kv: -kv[1]):
        if re.match(r"(LD|ST|FFMA|FADD|FMUL|DFMA|DADD|DMUL|UTMA|UBLK|LDG|STG|LDS|STS|LDC|SHFL|F2I|I2F|BAR)", op):
            lines.append("| %s | %d |" % (op, c))
    path = os.path.join(ROOT, "profiles", args.out + ".md")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(path, n, "instructions;", ", ".join("%s %d" % kv for kv in base.most_common(8)))


if __name__ == "__main__":
    main()
