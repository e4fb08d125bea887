#!/usr/bin/env python
"""Summarise an ncu report holding SEVERAL launches (ncu --set full -c N ...) into one CSV under profiles/: one column
per captured launch.  usage: ncu_multi_summary.py report.ncu-rep out.csv "comment line" """
import csv
import subprocess
import sys

rep, out, comment = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2:]
WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum launch__grid_size launch__block_size
launch__registers_per_thread launch__shared_mem_per_block_dynamic launch__waves_per_multiprocessor sm__cycles_active.avg
sm__cycles_active.max sm__inst_executed.avg.per_cycle_elapsed smsp__inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio sm__warps_active.avg.per_cycle_active
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum lts__t_sector_hit_rate.pct""".split()
WANT += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"))
with open(out, "w") as f:
    f.write("# %s\n" % comment)
    f.write("metric,unit," + ",".join("launch%d" % i for i in range(len(vals))) + "\n")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            f.write("%s,%s,%s\n" % (w, units[i], ",".join(v[i] for v in vals)))
print(open(out).read())
