#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` capture for profiles/:

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv
    python tools/ncu_summarize.py gpurun_out/prof_raw.csv [--traffic profiles/traffic.json] > profiles/rN_ncu_summary.md

One row per kernel (launches averaged): duration, DRAM bytes read / written, registers, grid, block, active warps,
DRAM throughput as a fraction of peak, shared-memory configuration.  Durations under the profiler are cold-cache and
serialised; the DRAM bytes are what a launch really moved (they feed `roofline.traffic` of the bench line)."""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
COLS = {"dur_us": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "regs": "launch__registers_per_thread", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smem_cfg": "launch__shared_mem_config_size"}


def short(name):
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name).strip()


def main():
    args = sys.argv[1:]
    traffic_path = None
    if "--traffic" in args:
        i = args.index("--traffic")
        traffic_path = args[i + 1]
        del args[i:i + 2]
    rows = [r for r in csv.reader(open(args[0])) if r]
    start = next(i for i, r in enumerate(rows) if r[0] == "ID")
    header, units, data = rows[start], rows[start + 1], rows[start + 2:]
    col = {}
    for key, metric in COLS.items():
        idx = [i for i, h in enumerate(header) if h == metric] or [i for i, h in enumerate(header) if h.endswith("." + metric)]
        col[key] = idx[0] if idx else None
    kname, kblock, kgrid = header.index("Kernel Name"), header.index("Block Size"), header.index("Grid Size")

    def val(r, key):
        i = col[key]
        if i is None or i >= len(r) or r[i] in ("", "no data", "n/a"):
            return None
        return float(r[i].replace(",", "")) * (UNIT.get(units[i], 1.0) if key in ("dur_us", "rd", "wr", "smem_cfg") else 1.0)

    kernels = {}
    for r in data:
        if len(r) <= kname:
            continue
        k = kernels.setdefault(short(r[kname]), {"n": 0, "grid": r[kgrid], "block": r[kblock], "sum": {}})
        k["n"] += 1
        for key in COLS:
            v = val(r, key)
            if v is not None:
                k["sum"][key] = k["sum"].get(key, 0.0) + v
    print("| kernel | launches | duration us | dram read B | dram write B | regs | grid | block | warps active % | dram % of peak | smem cfg B |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    traffic = {}
    for name, k in sorted(kernels.items(), key=lambda kv: -kv[1]["sum"].get("dur_us", 0.0) / kv[1]["n"]):
        m = {key: k["sum"][key] / k["n"] for key in k["sum"]}
        traffic[name] = m.get("rd", 0.0) + m.get("wr", 0.0)
        fmt = lambda key, spec: format(m[key], spec) if key in m else "-"
        print(f"| `{name}` | {k['n']} | {fmt('dur_us', '.2f')} | {fmt('rd', '.0f')} | {fmt('wr', '.0f')} | {fmt('regs', '.0f')} | "
              f"{k['grid']} | {k['block']} | {fmt('warps', '.1f')} | {fmt('dram_pct', '.2f')} | {fmt('smem_cfg', '.0f')} |")
    if traffic_path:
        json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
