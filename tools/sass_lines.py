#!/usr/bin/env python
"""Static SASS listing of one kernel with source lines, from the built library (no GPU needed).

  python tools/sass_lines.py <kernel name substring> [file.cuh first last]

Extracts the cubin (cuobjdump -xelf), disassembles it with line info (nvdisasm -g -c), and prints, for the chosen
kernel, either the instruction count per source line (default) or the instructions of a line range."""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.abspath(os.environ.get("RMD_LIB", os.path.join(ROOT, "rmdetect_b200", "librmdetect_b200.so")))
kern = sys.argv[1]
rng = (sys.argv[2], int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) >= 5 else None
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [f for f in glob.glob(tmp + "/*.cubin") if "ingest" not in f][0]
text = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
inside, cur, counts, listing = False, None, collections.Counter(), []
for ln in text:
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        counts[cur] += 1
        if rng and cur and cur[0] == rng[0] and rng[1] <= cur[1] <= rng[2]:
            listing.append("L%4d %s" % (cur[1], re.sub(r"\s+", " ", ln.strip())))
if rng:
    print("\n".join(listing))
    print("instructions in range:", len(listing))
else:
    for (f, l), n in sorted(counts.items(), key=lambda kv: (kv[0] or ("", 0))):
        print("%-22s %5d  %4d" % (f, l, n))
    print("total", sum(counts.values()))
