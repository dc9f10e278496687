#!/usr/bin/env python
"""Build profiles/ncu_traffic.json from `ncu --page raw --csv` exports of single-launch `--set full` captures
(tools/ncu_capture.sh writes gpurun_out/<tag>_raw.csv).

  python tools/ncu_traffic.py --config dh-32x4-b256 --source "..." name=gpurun_out/tag_raw.csv [name=...] > profiles/ncu_traffic.json

bench.py reads the file to fill roofline.traffic (labelled from_profile) when the configuration prefix matches.
"""
import argparse
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(name, scale=True):
        i = col.get(name)
        if i is None or vals[i] in ("", "n/a"):
            return None
        v = float(vals[i].replace(",", ""))
        return v * UNIT.get(units[i], 1.0) if scale else v

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    return {
        "kernel_name": vals[col["Kernel Name"]],
        "grid": vals[col["Grid Size"]],
        "duration_ms": get("gpu__time_duration.sum"),
        "dram_bytes_per_launch": (rd or 0.0) + (wr or 0.0),
        "dram_read": rd,
        "dram_write": wr,
        "dram_throughput_pct": next((float(vals[i]) for k, i in col.items() if k.endswith("dram__throughput.avg.pct_of_peak_sustained_elapsed")
                                     and vals[i] not in ("", "n/a")), None),
        "fmaheavy_pct": get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", False),
        "alu_pct": get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", False),
        "fp64_pct": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", False),
        "issue_active": get("smsp__issue_active.avg.per_cycle_active", False),
        "warps_active_per_sm": get("sm__warps_active.avg.per_cycle_active", False),
        "registers": get("launch__registers_per_thread", False),
        "inst_executed": get("smsp__inst_executed.sum", False),
        "stalls_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(vals[i])
                             for k, i in col.items() if "issue_stalled_" in k and k.endswith("per_issue_active.ratio")
                             and vals[i] not in ("", "n/a") and float(vals[i]) >= 0.3 and "selected" not in k},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--source", default="")
    ap.add_argument("pairs", nargs="+")
    a = ap.parse_args()
    out = {"config": a.config, "source": a.source, "kernels": {}}
    for p in a.pairs:
        name, path = p.split("=", 1)
        out["kernels"][name] = load(path)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
