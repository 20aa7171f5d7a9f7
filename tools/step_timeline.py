"""print the per-kernel device time of one training step from an ncu --metrics gpu__time_duration.sum csv"""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    seq.append((row["Kernel Name"][:46], v, row.get("Grid Size", "")))
idx = [i for i, (n, _, _) in enumerate(seq) if "exb_pull" in n]
tot = 0
for n, v, g in seq[idx[0]:idx[1]]:
    print("%8.1f us  %-46s %s" % (v, n, g))
    tot += v
print("step total %.1f us over %d launches" % (tot, idx[1] - idx[0]))
