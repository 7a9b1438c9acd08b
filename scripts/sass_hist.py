#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of libsgc_b200.so (cuobjdump --dump-sass).

usage: sass_hist.py <substring of the mangled kernel name> [path/to/lib.so]
Prints: instruction count, opcode histogram (IMAD variants kept apart: the FMA-heavy pipe also
executes IMAD.MOV / IMAD.IADD / IMAD.X), local-memory spill instructions, and the backward branches
(loops) with the number of instructions they span."""
import collections
import re
import subprocess
import sys


def functions(so):
    out = subprocess.run(["cuobjdump", "--dump-sass", so], capture_output=True, text=True).stdout
    name, body, res = None, [], {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                res[name] = body
            name, body = m.group(1), []
        elif name is not None:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        res[name] = body
    return res


def opcode(text):
    t = re.sub(r"^@!?U?P[0-9T]+\s+", "", text)
    op = t.split()[0]
    m = re.match(r"(IMAD\.(?:MOV|WIDE|X|IADD|SHL|HI))", op)
    return m.group(1) if m else op.split(".")[0]


def main():
    pat = sys.argv[1]
    so = sys.argv[2] if len(sys.argv) > 2 else "spacegraphcats_b200/libsgc_b200.so"
    fns = functions(so)
    hits = [n for n in fns if pat in n]
    if len(hits) != 1:
        print("kernels matching %r: %s" % (pat, hits or list(fns)[:40]))
        return 1
    body = fns[hits[0]]
    print("kernel %s: %d SASS instructions" % (hits[0], len(body)))
    hist = collections.Counter(opcode(t) for _, t in body)
    for op, n in hist.most_common():
        print("%6d  %s" % (n, op))
    print("local-memory (spill) instructions: %d" % sum(hist[o] for o in ("LDL", "STL")))
    addr_index = {a: i for i, (a, _) in enumerate(body)}
    for i, (a, t) in enumerate(body):
        m = re.search(r"\bBRA\S*\s+.*?(0x[0-9a-f]+)", t)
        if m and "BRA" in t:
            tgt = int(m.group(1), 16)
            if tgt in addr_index and addr_index[tgt] <= i:
                j = addr_index[tgt]
                span = body[j:i + 1]
                sp = sum(1 for _, x in span if re.search(r"\b(LDL|STL)\b", x))
                print("loop: instructions %d..%d (%d), spill instr inside: %d" % (j, i, i - j + 1, sp))
    return 0


if __name__ == "__main__":
    sys.exit(main())
