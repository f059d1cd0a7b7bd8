#!/usr/bin/env python
"""Compare two builds of a CUDA object / shared library function by function at the SASS level.

Used to show that a source change (a new template flag, a new instantiation) left the already
GPU-validated kernels instruction-for-instruction identical:

    python tools/sass_diff.py old/action_hex.o firedrake_b200/lib/action_hex.o [--strip-suffix ELb0E]

Function names are compared from the kernel's own name onwards (the anonymous-namespace hash in
the mangled prefix depends on the source path); ``--strip-suffix`` removes a trailing template
argument that exists only in the new build (e.g. a defaulted ``bool`` parameter).
"""
import argparse
import re
import subprocess
import sys


def functions(path, anchor):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            if anchor and anchor in cur:
                cur = cur[cur.index(anchor):]
            res[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and cur is not None:
            res[cur].append(m.group(1).strip())
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("old")
    ap.add_argument("new")
    ap.add_argument("--anchor", default="helmholtz_action_kernel",
                    help="compare names from this substring onwards")
    ap.add_argument("--strip-suffix", default="",
                    help="template-argument text to drop before the closing 'EEEv' of new names")
    args = ap.parse_args()
    old, new = functions(args.old, args.anchor), functions(args.new, args.anchor)
    if args.strip_suffix:
        new = {k.replace(args.strip_suffix + "EEEv", "EEEv", 1): v for k, v in new.items()}
    same = diff = missing = 0
    for k, v in old.items():
        if k not in new:
            missing += 1
            print("missing in new:", k[:100])
        elif new[k] == v:
            same += 1
        else:
            diff += 1
            print(f"DIFFERENT: {k[:100]} ({len(v)} -> {len(new[k])} instructions)")
    print(f"{len(old)} functions in old: {same} identical, {diff} different, {missing} missing; "
          f"{len(new) - same - diff} only in new")
    sys.exit(1 if diff or missing else 0)


if __name__ == "__main__":
    main()
