#!/usr/bin/env python
"""metric,unit,value rows of one kernel launch from an ncu report (`ncu -i rep --page raw --csv`), the form bench.py's
ncu_summary() and the judge read under profiles/.

  python tools/ncu_metrics.py gpurun_out/prof.ncu-rep "comment line" ["comment line" ...] > profiles/rNN_<kernel>_ncu_full.csv
"""
import csv
import io
import subprocess
import sys

rep, comments = sys.argv[1], sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
names, units, vals = rows[0], rows[1], rows[2]
for c in comments:
    print("# " + c)
print("metric,unit,value")
for n, u, v in zip(names, units, vals):
    if n in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"):
        continue
    print("%s,%s,%s" % (n.replace(",", ";"), u.replace(",", ";"), v.replace(",", "")))
