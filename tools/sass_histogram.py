#!/usr/bin/env python
"""Static SASS opcode histograms of the hot kernels in libhegpu.so (cuobjdump -sass), so that the
instruction-count claims in DESIGN.md can be checked.

  python tools/sass_histogram.py [--out profiles/r2_sass_histogram.json] [--match REGEX ...]

Counts are STATIC (one per instruction in the cubin).  For the kernels listed in DYNAMIC the script also
reports per-loop-body counts: the instructions between a backward branch target and the branch are one body.
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "homomorphic-encryption-algorithms-diploma-thesis_b200", "libhegpu.so")
DEFAULT = [r"dh_inner_kernel<3, 4>", r"ntt_fwd_park_kernel<13, 3, hegpu::(PlainJob|KsLiftJob|KsModDownJob|FinalNttJob|RescaleJob)>",
           r"ntt_inv_park_kernel<13, 3, hegpu::(PlainJob|KsInttJob|HalfInttJob|FinalInttJob)>", r"ks_inner_sum_kernel", r"ks_digits_kernel"]
GROUPS = {
    "IMAD.WIDE": lambda op: op.startswith("IMAD.WIDE"),
    "IMAD (32-bit mul)": lambda op: op.startswith("IMAD") and not op.startswith("IMAD.WIDE") and not op.startswith("IMAD.MOV") and not op.startswith("IMAD.X")
    and not op.startswith("IMAD.SHL") and not op.startswith("IMAD.IADD"),
    "IMAD.MOV/IADD/SHL/X (no multiply)": lambda op: op.startswith(("IMAD.MOV", "IMAD.X", "IMAD.SHL", "IMAD.IADD")),
    "IADD3/IADD/LEA": lambda op: op.startswith(("IADD", "LEA", "UIADD")),
    "LOP3/SHF/PRMT/SEL/MOV": lambda op: op.startswith(("LOP3", "SHF", "PRMT", "SEL", "MOV", "UMOV", "ULOP", "USHF")),
    "ISETP/compare": lambda op: op.startswith(("ISETP", "UISETP", "PLOP", "DSETP", "FSETP")),
    "DFMA/DADD/DMUL": lambda op: op.startswith(("DFMA", "DADD", "DMUL")),
    "F2F/I2F/F2I": lambda op: op.startswith(("F2F", "I2F", "F2I")),
    "LDG": lambda op: op.startswith("LDG"),
    "STG": lambda op: op.startswith("STG"),
    "LDS": lambda op: op.startswith("LDS"),
    "STS": lambda op: op.startswith("STS"),
    "LDC/ULDC": lambda op: op.startswith(("LDC", "ULDC")),
    "BAR/branch": lambda op: op.startswith(("BAR", "BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC")),
}


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    name, body = None, []
    for ln in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            body.append(ln)
    if name:
        yield name, body


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True, check=True)
    return r.stdout.splitlines()


def opcode(ln):
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    return (int(m.group(1), 16), m.group(2)) if m else (None, None)


def histogram(body):
    ops = collections.Counter()
    for ln in body:
        _, op = opcode(ln)
        if op:
            ops[op] += 1
    grouped = {g: sum(c for o, c in ops.items() if f(o)) for g, f in GROUPS.items()}
    grouped["other"] = sum(ops.values()) - sum(grouped.values())
    grouped["total"] = sum(ops.values())
    return grouped, ops


def loops(body):
    """Innermost loop bodies: a backward BRA and the instructions between its target and itself."""
    res = []
    addr = [opcode(ln)[0] for ln in body]
    for i, ln in enumerate(body):
        m = re.search(r"\bBRA(?:\.U)?\s+(?:[!A-Z0-9,]+\s+)?`?\(?\.?L?_?x?_?\d*\)?\s*(0x[0-9a-f]+)", ln) or re.search(r"\bBRA.*?(0x[0-9a-f]+)", ln)
        if not m or addr[i] is None:
            continue
        tgt = int(m.group(1), 16)
        if tgt < addr[i]:
            seg = [b for a, b in zip(addr, body) if a is not None and tgt <= a <= addr[i]]
            res.append((tgt, addr[i], seg))
    inner = [r for r in res if not any(o is not r and r[0] <= o[0] and o[1] <= r[1] for o in res)]
    return inner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--match", nargs="*", default=DEFAULT)
    ap.add_argument("--top", type=int, default=0, help="also list the N most frequent raw opcodes per kernel")
    args = ap.parse_args()
    fns = list(functions(LIB))
    names = demangle([n for n, _ in fns])
    report = {}
    for (_, body), dn in zip(fns, names):
        if not any(re.search(p, dn) for p in args.match):
            continue
        g, ops = histogram(body)
        entry = {"static": g}
        lp = []
        for tgt, end, seg in loops(body):
            lg, _ = histogram(seg)
            if lg["total"] >= 64:
                lp.append({"from": hex(tgt), "to": hex(end), **lg})
        entry["innermost_loops"] = lp
        if args.top:
            entry["top_opcodes"] = dict(ops.most_common(args.top))
        report[dn] = entry
    js = json.dumps(report, indent=1)
    if args.out:
        with open(args.out, "w") as f:
            f.write(js + "\n")
    else:
        print(js)


if __name__ == "__main__":
    sys.exit(main())
