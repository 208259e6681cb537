"""Regenerate profiles/ncu_traffic.json from ncu captures of the SHIPPED build.

    python tools/ncu_traffic.py            # reads profiles/r2_ncu_traffic_*.csv, writes profiles/ncu_traffic.json

Each input is `ncu -i <rep> --page raw --csv` of one capture made by tools/gpu_session_r2c.sh:
    r2_ncu_traffic_n1.csv      bench.py --gpus 1 (1 -> 1, 16 GB payload)
    r2_ncu_traffic_x2_n{2,4,8}.csv   2-GPU emulation of dest rank 0 of the FSDP(N)->TP(N) sync
                                     (tools/sweep_plan.py --mode x2: the real rect tables and the real
                                     local/NVLink byte mix of one rank; ncu serialises the two GPUs, so the
                                     peer is idle while the profiled kernel runs)
bench.py reads the result for `roofline.traffic`; nobody edits the JSON by hand.
"""

from __future__ import annotations

import csv
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "": 1, "%": 1}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            try:
                d[h] = float(v.replace(",", "")) * UNIT.get(u, 1)
            except ValueError:
                d[h] = v
        out.append(d)
    return out


def main():
    table = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_ncu_traffic_*.csv"))):
        m = re.search(r"r2_ncu_traffic_(?:x2_)?n(\d+)\.csv$", path)
        if not m:
            continue
        n = m.group(1)
        rows = [r for r in load(path) if "copy_rects" in str(r.get("Kernel Name", ""))]
        if not rows:
            continue
        r = rows[0]
        dram = r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
        table[n] = {
            "dram_bytes": dram,
            "dram_read_bytes": r.get("dram__bytes_read.sum"),
            "dram_write_bytes": r.get("dram__bytes_write.sum"),
            "nvlrx_bytes": r.get("nvlrx__bytes.sum"),
            "nvlrx_user_bytes": r.get("nvlrx__bytes_data_user.sum"),
            "nvltx_bytes": r.get("nvltx__bytes.sum"),
            "kernel_time_s_under_ncu": r.get("gpu__time_duration.sum"),
            "grid": r.get("launch__grid_size"),
            "block": r.get("launch__block_size"),
            "registers": r.get("launch__registers_per_thread"),
            "source": os.path.basename(path) + (" (2-GPU emulation of one rank's plan; peer idle under ncu)" if "x2" in path else ""),
        }
    out = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    with open(out, "w") as f:
        json.dump(table, f, indent=1)
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
