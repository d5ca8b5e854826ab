"""Summarise an .ncu-rep (ncu --set full) per launch: duration, DRAM bytes and GB/s, tensor-pipe %, occupancy.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep [max_rows]"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct"]
TO_B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}

def main(path, limit=60):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {k: hdr.index(k) for k in KEYS if k in hdr}
    print(f"# {path}: ncu --set full --clock-control none (per launch; DRAM GB/s = (read+write)/duration)")
    print(f"{'kernel':44s} {'grid':>14s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'dram%':>6s} {'tensor%':>7s} {'sm%':>5s} {'occ%':>5s} {'regs':>4s}")
    for r in rows[2:2 + limit]:
        name = r[hdr.index("Kernel Name")].replace("void ", "").split("(")[0][:44]
        def val(k, conv=None):
            if k not in ix: return float("nan")
            v = float(r[ix[k]].replace(",", ""))
            return v * conv.get(units[ix[k]], 1.0) if conv else v
        us = val("gpu__time_duration.sum", TO_US)
        rd, wr = val("dram__bytes_read.sum", TO_B), val("dram__bytes_write.sum", TO_B)
        print(f"{name:44s} {r[hdr.index('Grid Size')]:>14s} {us:8.1f} {rd/1e6:8.1f} {wr/1e6:8.1f} {(rd+wr)/us/1e3:7.0f} "
              f"{val('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{val('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} "
              f"{val('sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
              f"{val('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} {val('launch__registers_per_thread'):4.0f}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
