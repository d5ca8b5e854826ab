"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
   python tools/summarize_launches.py gpurun_out/launches.csv [top_n]"""
import csv, re, sys
from collections import defaultdict

def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
        ms = v * scale
        name = re.sub(r"<.*", "", name)[:90]
        agg[name][0] += 1
        agg[name][1] += ms
        total += ms
    print(f"{sum(a[0] for a in agg.values())} launches, {total:.2f} ms total device time (serialised, cold cache)")
    print(f"{'kernel':90s} {'calls':>6s} {'ms':>9s} {'share':>6s}")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{name:90s} {n:6d} {ms:9.3f} {100*ms/total:5.1f}%")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
