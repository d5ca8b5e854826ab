"""Summarise a device timeline written by `bench.py --timeline FILE` (two steps of rank 0, CUPTI through torch.profiler):
per step the compute stream's span, busy time and idle gaps, and every NCCL kernel with its start and duration.
   python tools/timeline_summary.py profiles/r02_timeline_4gpu.csv [gap_us]
The numbers quoted in DESIGN.md section 8 / profiles/r02_summary.md come from this script."""
import csv
import sys


def main(path, gap_us=15.0):
    ev = [(float(r["start_us"]), float(r["dur_us"]), r["stream"], r["name"]) for r in csv.DictReader(open(path))]
    streams = {}
    for e in ev:
        streams.setdefault(e[2], []).append(e)
    main_id = max(streams, key=lambda s: len(streams[s]))            # the compute stream carries almost every kernel
    main = streams[main_id]
    print(f"{path}: {len(ev)} device activities on streams {sorted(streams)}; compute stream = {main_id}")
    half = len(main) // 2
    for k, step in enumerate((main[:half], main[half:])):
        busy = sum(e[1] for e in step)
        span = step[-1][0] + step[-1][1] - step[0][0]
        gaps = [(b[0] - (a[0] + a[1]), a[0]) for a, b in zip(step[:-1], step[1:])]
        big = [(round(g), round(t / 1e3, 2)) for g, t in gaps if g > gap_us]
        small = sum(g for g, _ in gaps if 0 < g <= gap_us)
        print(f"step {k}: {len(step)} activities, span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, "
              f"gaps <= {gap_us:.0f} us total {small / 1e3:.2f} ms, larger gaps (us, at ms): {big}")
    print("NCCL kernels (start ms, duration us, stream):")
    for e in ev:
        if "nccl" in e[3].lower():
            print(f"  {e[0] / 1e3:9.3f} {e[1]:9.1f}  {e[2]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 15.0)
