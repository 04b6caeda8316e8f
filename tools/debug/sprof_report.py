#!/usr/bin/env python3
"""Report of tools/debug/sprof.c (a SIGPROF sampling profiler for boxes without perf: LD_PRELOAD it, it writes the sampled
instruction pointers and the process's memory map on exit): samples per function and per source line of one shared object.
usage:  gcc -O2 -shared -fPIC -o /tmp/sprof.so tools/debug/sprof.c
        (build the library with -g)  LD_PRELOAD=/tmp/sprof.so SPROF_OUT=/tmp/sprof.out KNASTER_GPU_LIB=<lib> python tools/host_bench.py subtractive 16384 10 80 1
        tools/debug/sprof_report.py /tmp/sprof.out <lib>"""
import collections
import re
import subprocess
import sys

out_file, lib = sys.argv[1], sys.argv[2]
name = lib.split("/")[-1]
maps, rips = [], []
for line in open(out_file):
    if line.startswith("M "):
        m = re.match(r"M ([0-9a-f]+)-([0-9a-f]+) (\S+) ([0-9a-f]+) \S+ \S+\s+(\S+)", line)
        if m:
            maps.append((int(m.group(1), 16), int(m.group(2), 16), int(m.group(4), 16), m.group(5)))
    else:
        rips.append(int(line, 16))
offs, elsewhere = [], collections.Counter()
for r in rips:
    for a, b, off, n in maps:
        if a <= r < b:
            if name in n:
                offs.append(r - a + off)
            else:
                elsewhere[n.split("/")[-1]] += 1
            break
    else:
        elsewhere["(unmapped: the interpreter, anonymous memory)"] += 1
print(f"{len(rips)} samples, {len(offs)} in {name}; elsewhere: {elsewhere.most_common(4)}")
res = subprocess.run(["addr2line", "-e", lib, "-f", "-C"] + [hex(o) for o in offs], capture_output=True, text=True).stdout.splitlines()
fn, ln = collections.Counter(), collections.Counter()
for i in range(0, len(res), 2):
    fn[res[i][:70]] += 1
    ln[res[i + 1].split("/")[-1].split(" ")[0]] += 1
print("--- functions")
for k, v in fn.most_common(12):
    print(f"{v:6d}  {k}")
print("--- source lines")
for k, v in ln.most_common(25):
    print(f"{v:6d}  {k}")
