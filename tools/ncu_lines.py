#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i rep --page source --csv) with nvdisasm line info so that
executed-instruction counts and stall samples can be read per CUDA source line.
usage: ncu_lines.py src.csv kernel.cubin [top_n]"""
import collections
import csv
import re
import subprocess
import sys

src_csv, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
kfilter = sys.argv[4] if len(sys.argv) > 4 else "step_kernel"
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
in_k, line, order = False, None, []
for ln in dis:
    if ".text." in ln:
        in_k = kfilter in ln
    if not in_k:
        continue
    m = re.search(r'//## File ".*?", line (\d+)(.*)', ln)
    if m:
        line = int(m.group(1))
        inl = re.findall(r'inlined at ".*?", line (\d+)', m.group(2))
        outer = int(inl[-1]) if inl else line
        cur = (line, outer)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        order.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
isrc = hdr.index("Source")
body = rows[2:]
if len(body) != len(order):
    print("warning: %d sass rows vs %d disassembled instructions" % (len(body), len(order)))
by_line = collections.Counter(); by_outer = collections.Counter(); samp_outer = collections.Counter(); samp_line = collections.Counter()
tot = 0
for r, (l, o) in zip(body, order):
    n = int(r[ii] or 0); s = int(r[isamp] or 0)
    by_line[l] += n; by_outer[o] += n; samp_outer[o] += s; samp_line[l] += s; tot += n
tots = sum(samp_outer.values())
print("total warp instructions", tot, "samples", tots)
print("--- by outermost (kernel-body) line: inst%  samples%")
for l, n in by_outer.most_common(top):
    print("line %5d  %6.2f%%  %6.2f%%" % (l, 100.0 * n / tot, 100.0 * samp_outer[l] / max(tots, 1)))
print("--- by innermost line")
for l, n in by_line.most_common(top):
    print("line %5d  %6.2f%%  %6.2f%%" % (l, 100.0 * n / tot, 100.0 * samp_line[l] / max(tots, 1)))
