"""Condense an `ncu --csv` launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum)
into one row per launch of the LAST V-cycle in the log: kernel, grid, block, time, DRAM MB, share of the cycle.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'^k_(st|fix|coarse_gemv|jacobi|residual|prolong|colour)' --csv --log-file gpurun_out/launches.csv \
        python tools/gpu_probe.py --cycles 2
    python tools/ncu_launches.py gpurun_out/launches.csv 4 > profiles/<name>.csv \
        [--traffic profiles/traffic.json --shape 512x512x512 --smoother jacobi|rbgs]

The second argument is the number of V-cycles in the log: `gpu_probe.py --cycles 2` runs two warm-up cycles (one per
ping-pong parity, omg_bench_cycles) and the two timed ones, so 4; the last cycle is the one condensed.
Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py, not absolutes.
"""
import csv
import json
import sys


def main():
    path, ncycles = sys.argv[1], int(sys.argv[2])
    traffic_out = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    per = {}
    order = []
    for r in rd:
        i = int(r["ID"])
        if i not in per:
            per[i] = {"kernel": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]}
            order.append(i)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name == "gpu__time_duration.sum":
            per[i]["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        else:
            mb = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
            per[i]["rd" if "read" in name else "wr"] = mb
    n = len(order) // ncycles
    last = order[-n:]
    total = sum(per[i]["us"] for i in last)
    w = csv.writer(sys.stdout)
    w.writerow(["id", "kernel", "grid", "block", "time_us", "dram_read_MB", "dram_write_MB", "share_of_cycle"])
    for i in last:
        p = per[i]
        w.writerow([i, p["kernel"], p["grid"], p["block"], "%.2f" % p["us"], "%.2f" % p.get("rd", 0.0),
                    "%.2f" % p.get("wr", 0.0), "%.4f" % (p["us"] / total)])
        rows.append(p)
    w.writerow(["", "cycle total", "", "", "%.2f" % total, "%.2f" % sum(p.get("rd", 0) for p in rows),
                "%.2f" % sum(p.get("wr", 0) for p in rows), "1.0"])
    if traffic_out:
        # the three level-0 kernels of a V(1,1) cycle are the first two and the last launch; bench.py looks the
        # dominant kernel up under the workload shape ({"<shape>": {"<name>@L0": DRAM bytes per launch}, "source": ..})
        shape = sys.argv[sys.argv.index("--shape") + 1] if "--shape" in sys.argv else "512x512x512"
        smoother = sys.argv[sys.argv.index("--smoother") + 1] if "--smoother" in sys.argv else "jacobi"
        # (Jacobi: the pre-smoothing sweep and the restricted residual are one launch, k_jr3)
        names = {"jacobi": ["jacobi_residual_restrict@L0", "prolong_jacobi@L0"],
                 "rbgs": ["rbgs_sweep@L0", "residual_restrict@L0", "prolong_rbgs_sweep@L0"]}[smoother]
        sel = [rows[0], rows[-1]] if smoother == "jacobi" else [rows[0], rows[1], rows[-1]]
        try:
            out = json.load(open(traffic_out))
            if not isinstance(out.get(shape, {}), dict) or "source" not in out:
                out = {}
        except Exception:  # noqa: BLE001
            out = {}
        out.setdefault(shape, {}).update({k: (p.get("rd", 0) + p.get("wr", 0)) * 1e6 for k, p in zip(names, sel)})
        out["source"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu launch lists "
                         "profiles/r2_launches_*.csv (python tools/gpu_probe.py --cycles 2), not measured in the timed run")
        json.dump(out, open(traffic_out, "w"), indent=1)


if __name__ == "__main__":
    main()
