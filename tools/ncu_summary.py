"""Condenses an `ncu --set full` report into the per-kernel table kept under profiles/ and (optionally) refreshes
profiles/traffic.json (DRAM bytes per launch of the composites, read by bench.py's roofline object).

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
  python tools/ncu_summary.py /tmp/raw.csv profiles/r01g_ncu_full_summary.csv [--traffic C2]
"""
import csv
import json
import os
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def to_bytes(v, unit):
    m = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1.0)


def main():
    raw, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(c) if c in hdr else None for c in COLS]
    kn = hdr.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{c} [{units[i]}]" if i is not None else c for c, i in zip(COLS, idx)])
        for d in data:
            name = d[kn].split("(")[0].replace("void ", "").replace("b2r::", "")
            w.writerow([name] + [d[i] if i is not None else "" for i in idx])
    if "--traffic" in sys.argv:
        wl = sys.argv[sys.argv.index("--traffic") + 1]
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        path = os.path.join(root, "profiles", "traffic.json")
        t = json.load(open(path)) if os.path.exists(path) else {}
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        acc = {}
        for d in data:
            for key in ("composite_fwd", "composite_bwd"):
                if key in d[kn]:
                    acc.setdefault(key, []).append(to_bytes(d[ir], units[ir]) + to_bytes(d[iw], units[iw]))
        t[wl] = {k: sum(v) / len(v) for k, v in acc.items()}
        t["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full (profiles/*_ncu_full_summary.csv); "
                         "mean over the captured launches")
        json.dump(t, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
