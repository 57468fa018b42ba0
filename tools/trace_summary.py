"""Summarises a chrome trace written by `bench.py --trace-e2e` or tools/lanes_timeline.py (GPU activities only)."""
import json
import sys

import numpy as np


def main(path):
    ev = json.load(open(path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
    t0 = min(e["ts"] for e in ks)
    t1 = max(e["ts"] + e["dur"] for e in ks)
    print(f"span {t1 - t0:.1f} us, {len(ks)} GPU activities")
    by = {}
    for e in ks:
        n = e["name"].split("(")[0].replace("void ", "").replace("b2r::", "")[:48]
        if e["cat"] == "gpu_memcpy":
            n = "memcpy " + e["name"][:20]
        by.setdefault(n, []).append(e["dur"])
    print(f"{'activity':50s} count   mean us      max    total")
    for n, d in sorted(by.items(), key=lambda kv: -sum(kv[1]))[:28]:
        print(f"{n:50s} {len(d):5d} {sum(d) / len(d):9.1f} {max(d):8.1f} {sum(d):8.1f}")
    for cat in ("kernel", "gpu_memcpy"):
        sel = [e for e in ks if e["cat"] == cat]
        if not sel:
            continue
        grid = np.linspace(t0, t1, 801)[:-1]
        conc = np.zeros(len(grid), int)
        for e in sel:
            conc += (grid >= e["ts"]) & (grid < e["ts"] + e["dur"])
        print(cat, "busy fraction of span:", round(float((conc > 0).mean()), 3), " mean concurrency:", round(float(conc.mean()), 2))
    cp = [e for e in ks if e["cat"] == "gpu_memcpy"]
    for kind in ("HtoD", "DtoH"):
        c = [e for e in cp if kind in e["name"]]
        if c:
            b = sum(e.get("args", {}).get("bytes", 0) for e in c)
            d = sum(e["dur"] for e in c)
            print(f"{kind}: {len(c)} copies, {b / 1e6:.1f} MB, busy {d:.0f} us, {b / max(d, 1e-9) / 1e3:.1f} GB/s while copying; "
                  f"first start {min(e['ts'] for e in c) - t0:.0f}, last end {max(e['ts'] + e['dur'] for e in c) - t0:.0f}")


if __name__ == "__main__":
    main(sys.argv[1])
