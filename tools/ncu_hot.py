"""Top warp-stall sample locations (SASS) of an ncu report: python tools/ncu_hot.py report.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
i_s, i_src, i_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = rows[2:]
num = lambda x: float(x.replace(",", "")) if x not in ("", "-") else 0.0
tot = sum(num(r[i_s]) for r in data)
print("total samples %d, SASS lines %d, instructions executed %d" % (tot, len(data), sum(num(r[i_ex]) for r in data)))
order = sorted(range(len(data)), key=lambda i: -num(data[i][i_s]))[:N]
for i in sorted(order):
    r = data[i]
    print("%5d %6.2f%%  exec %9d  line %4d  %s" % (num(r[i_s]), 100 * num(r[i_s]) / tot, num(r[i_ex]), i, r[i_src].strip()[:100]))
