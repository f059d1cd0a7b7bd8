#!/usr/bin/env python
"""Per-unit dynamic profile of a persistent kernel from a cuobjdump -sass listing of ONE function:
finds the outermost back-branch (main loop) and the largest inner back-branch loop inside it (weight = its trip
count), then prints instruction counts, fp64 counts and the sum of the encoded stall counts (the cycles ONE warp
needs per main-loop trip if nothing but fixed latencies delayed it).

    python tools/sass_loops.py listing.sass [inner_trip_count=4]"""
import collections
import re
import sys


def parse(path):
    lines = open(path).read().split("\n")
    ins = []
    i = 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.search(r"/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            hi = int(m2.group(1), 16) if m2 else 0
            text = re.sub(r"^@!?U?P\w+\s+", "", m.group(2).strip())
            ins.append((int(m.group(1), 16), text, (hi >> 41) & 0xf))
            i += 2
        else:
            i += 1
    return ins


def main():
    ins = parse(sys.argv[1])
    trips = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    loops = []
    for a, t, _ in ins:
        m = re.match(r"BRA(?!\.DIV)\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    outer = max(loops, key=lambda l: l[1] - l[0])
    inner = max((l for l in loops if outer[0] <= l[0] and l[1] < outer[1] and l != outer),
                key=lambda l: l[1] - l[0], default=None)
    fp = {"DFMA", "DMUL", "DADD", "MUFU"}
    n = f64 = stall = 0.0
    ops = collections.Counter()
    for a, t, st in ins:
        if not (outer[0] <= a <= outer[1]):
            continue
        w = trips if inner and inner[0] <= a <= inner[1] else 1.0
        op = t.split()[0].split(".")[0]
        n += w
        stall += w * st
        ops[op] += w
        if op in fp:
            f64 += w
    print(f"main loop {outer[0]:#x}..{outer[1]:#x}" + (f", inner {inner[0]:#x}..{inner[1]:#x} x{trips:g}" if inner else ""))
    print(f"instructions {n:.0f}  fp64 {f64:.0f}  other {n - f64:.0f}  sum of stall counts {stall:.0f}")
    print("  ".join(f"{k}:{v:.0f}" for k, v in ops.most_common(14)))


if __name__ == "__main__":
    main()
