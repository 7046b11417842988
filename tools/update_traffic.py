"""Summarise one `ncu --set full` capture of the dominant kernel into
profiles/normals_kernel_traffic.json (what bench.py reports as roofline.traffic), stamped with a
hash of the kernel sources so that bench.py can tell a stale capture from a current one.

    python tools/update_traffic.py gpurun_out/<capture>.ncu-rep [note]
"""
import csv
import datetime
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "point-cloud-processing_b200", "csrc")


def csrc_hash():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and f.endswith((".cu", ".cuh", ".hpp", ".inc")):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, unit, val = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(head)}

    def num(name):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1.0, "us": 1e-3,
                 "%": 1.0, "inst": 1.0}.get(unit[col[name]], 1.0)
        return float(val[col[name]].replace(",", "")) * scale

    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    rec = {
        "kernel": val[col["Kernel Name"]],
        "source": os.path.basename(rep) + " (ncu --set full --clock-control none, one launch, "
                  "10M-point noisy plane, k=15)",
        "captured_at": datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%MZ"),
        "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"],
                                 capture_output=True, text=True).stdout.strip(),
        "csrc_sha16": csrc_hash(),
        "dram_bytes_per_launch": rd + wr,
        "dram_bytes_read": rd,
        "dram_bytes_write": wr,
        "duration_ms_under_ncu": num("gpu__time_duration.sum"),
        "l2_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
        "warp_instructions": num("smsp__inst_executed.sum"),
        "algorithmic_bytes_per_launch": 2040000000,
        "note": sys.argv[2] if len(sys.argv) > 2 else "",
    }
    with open(os.path.join(ROOT, "profiles", "normals_kernel_traffic.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
