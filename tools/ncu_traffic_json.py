"""DRAM traffic of one kernel from an `ncu --set full` report -> the small JSON bench.py quotes as roofline.traffic.

    python tools/ncu_traffic_json.py report.ncu-rep kernel_regex "command line" > profiles/rNN_<kernel>_traffic.json
"""
import csv
import json
import re
import subprocess
import sys


def main(path, pattern, command):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    launches, total, dur_us = 0, 0.0, 0.0
    for r in rows[2:]:
        if not re.search(pattern, r[col["Kernel Name"]]):
            continue
        launches += 1
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(r[col[k]].replace(",", "")) * scale[units[col[k]]]
        dur_us += float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[col["gpu__time_duration.sum"]], 1.0)
    print(json.dumps({"kernel": pattern, "launches": launches, "dram_bytes_per_launch": total / max(launches, 1),
                      "dram_bytes_total": total, "ncu_duration_us_total": dur_us, "command": command}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
