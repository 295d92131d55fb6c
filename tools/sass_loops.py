#!/usr/bin/env python3
"""Instruction mix of the inner loops of a kernel, from `cuobjdump -sass` text (a CPU-side check before GPU time).
usage: cuobjdump -sass lib.so | tools/sass_loops.py <kernel-name-substring> [min_shfl]
Prints every backward-branch loop that holds at least `min_shfl` SHFL.IDX (the rotation loops of k_force_sym)."""
import re, sys, collections
name = sys.argv[1]; min_shfl = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ins = []; on = False
for line in sys.stdin:
    if "Function :" in line:
        on = name in line
        if on: ins = []; cur = line.strip()
        continue
    if not on: continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: k for k, (a, _) in enumerate(ins)}
for k, (a, t) in enumerate(ins):
    m = re.search(r"BRA\s+(?:`\(\S+\)|0x([0-9a-f]+))", t)
    if m and m.group(1):
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            body = [x for _, x in ins[addr[tgt]:k + 1]]
            ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for x in body)
            if sum(v for o, v in ops.items() if o == "SHFL") >= min_shfl:
                print(f"loop 0x{tgt:x}..0x{a:x}: {len(body)} instructions: " + ", ".join(f"{o} {v}" for o, v in ops.most_common()))
