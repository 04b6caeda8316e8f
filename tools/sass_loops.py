#!/usr/bin/env python3
"""List the loops (backward branches) of one kernel in a cuobjdump -sass listing with their size and
instruction mix -- used to check what a source change did to a hot loop before spending GPU time.
usage: cuobjdump -sass x.o | tools/sass_loops.py <kernel-substring> [min_instructions]"""
import re, sys, collections
pat = sys.argv[1]; minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cur = None; ins = []
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m: ins.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r'\bBRA(?:\.\w+)* .*?0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr and i - addr[tgt] + 1 >= minlen:
            body = ins[addr[tgt]:i + 1]
            c = collections.Counter(re.sub(r'^@!?U?P\w+\s+', '', x[1]).split()[0].split('.')[0] for x in body)
            print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr  " + " ".join(f"{k}:{v}" for k, v in c.most_common(14)))
