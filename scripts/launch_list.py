"""Prints the last proof's kernels from an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python scripts/launch_list.py gpurun_out/x_launches.csv [n_last]"""
import csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = 0.0
for row in rows[-n:]:
    ms = float(row["Metric Value"].replace(",", "")) / (1e6 if row["Metric Unit"] == "ns" else 1e3 if row["Metric Unit"] in ("us", "usecond") else 1)
    tot += ms
    print("%-64s %-14s %-12s %8.4f" % (row["Kernel Name"][:64], row["Grid Size"], row["Block Size"], ms))
print("total %.3f ms over %d launches (of %d in the file)" % (tot, min(n, len(rows)), len(rows)))
