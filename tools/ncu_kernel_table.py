"""Turns an `ncu --csv --metrics ...` log (one row per launch and metric) into the per-kernel table kept under profiles/:
launches, total device time, and for each kernel's longest launch DRAM GB/s, L2 / L1 hit rates, issue-slot and warp
occupancy, L1 data-stage and L2 throughput, registers.
    python tools/ncu_kernel_table.py gpurun_out/allk.csv > profiles/rN_all_kernels_metrics.txt"""
import collections
import csv
import sys

M = {"t": "gpu__time_duration.sum", "dr": "dram__bytes_read.sum", "dw": "dram__bytes_write.sum",
     "l2": "lts__t_sector_hit_rate.pct", "l1": "l1tex__t_sector_hit_rate.pct",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "l1tex": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "regs": "launch__registers_per_thread"}
UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    idc, kn, mn, mu, mv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[idc], {"name": r[kn].split("(")[0].replace("void ", "").replace("tmk::", "")})
        v = float(r[mv].replace(",", "")) if r[mv] not in ("", "n/a") else 0.0
        d[r[mn]] = v * UNIT.get(r[mu], 1.0)
    per = collections.OrderedDict()
    for d in launches.values():
        k = per.setdefault(d["name"], {"n": 0, "total": 0.0, "best": None})
        k["n"] += 1
        k["total"] += d.get(M["t"], 0.0)
        if k["best"] is None or d.get(M["t"], 0.0) > k["best"].get(M["t"], 0.0):
            k["best"] = d
    print("%-46s %3s %9s %8s %9s %7s %7s %7s %7s %7s %6s %5s" % ("kernel", "n", "total_ms", "max_ms", "dramGB/s", "L2hit%", "L1hit%",
                                                                     "issue%", "warps%", "l1tex%", "lts%", "regs"))
    for name, k in sorted(per.items(), key=lambda x: -x[1]["total"]):
        b = k["best"]
        t = b.get(M["t"], 0.0)
        gbs = (b.get(M["dr"], 0.0) + b.get(M["dw"], 0.0)) / (t * 1e-3) / 1e9 if t else 0.0
        print("%-46s %3d %9.3f %8.3f %9.1f %7.1f %7.1f %7.1f %7.1f %7.1f %6.1f %5d" % (
            name[:46], k["n"], k["total"], t, gbs, b.get(M["l2"], 0), b.get(M["l1"], 0), b.get(M["issue"], 0), b.get(M["warps"], 0),
            b.get(M["l1tex"], 0), b.get(M["lts"], 0), int(b.get(M["regs"], 0))))


if __name__ == "__main__":
    main(sys.argv[1])
