#!/usr/bin/env python3
"""Where do the warp samples of a kernel go?  Reads `ncu -i x.ncu-rep --page source --csv` (SASS view)
and prints samples / executed instructions per region, regions being cut at backward-branch loops of
at least MIN instructions (the straight-line groups) -- everything else is 'other'.
usage: tools/ncu_src_regions.py src.csv [MIN]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
MIN = int(sys.argv[2]) if len(sys.argv) > 2 else 150
hdr = rows[1]; body = rows[2:]
A, S, SA, EX = hdr.index('Address'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
addr = [int(r[A], 16) for r in body]; base = addr[0]
idx = {a: i for i, a in enumerate(addr)}  # branch targets are absolute addresses on this page
loops = []
for i, r in enumerate(body):
    m = re.search(r'\bBRA(?:\.\w+)* .*?0x([0-9a-f]+)', r[S])
    if m:
        t = int(m.group(1), 16)
        if t in idx and idx[t] <= i and i - idx[t] + 1 >= MIN: loops.append((idx[t], i))
# keep innermost (smallest) loops only
loops.sort(key=lambda l: l[1] - l[0])
used = [False] * len(body); regions = []
for lo, hi in loops:
    if any(used[lo:hi + 1]): continue
    for k in range(lo, hi + 1): used[k] = True
    regions.append((lo, hi))
tot_s = sum(int(r[SA] or 0) for r in body); tot_e = sum(int(r[EX] or 0) for r in body)
print(f"total: {tot_s} samples, {tot_e} warp instructions executed, {len(body)} SASS instructions")
for lo, hi in sorted(regions):
    s = sum(int(r[SA] or 0) for r in body[lo:hi + 1]); e = sum(int(r[EX] or 0) for r in body[lo:hi + 1])
    trips = int(body[hi][EX] or 0)
    print(f"loop @{addr[lo]-base:#x}: {hi-lo+1:4d} instr, {trips} trips, samples {100*s/tot_s:5.1f} %, executed {100*e/tot_e:5.1f} %, samples/instr {s/max(e,1)*1e3:.3f}e-3")
s = sum(int(r[SA] or 0) for k, r in enumerate(body) if not used[k]); e = sum(int(r[EX] or 0) for k, r in enumerate(body) if not used[k])
print(f"other: samples {100*s/tot_s:5.1f} %, executed {100*e/tot_e:5.1f} %")
