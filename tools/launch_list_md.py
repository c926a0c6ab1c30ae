"""Turns an `ncu --metrics gpu__time_duration.sum --csv` log into the per-kernel share table kept under profiles/.

    python tools/launch_list_md.py launches.csv "title" "command line" > profiles/rNN_launch_list.md
"""
import collections
import csv
import re
import sys


def main(path, title, command):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    k_name, k_metric, k_val, k_unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    total, per = 0.0, collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= k_val or r[k_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[k_val].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(r[k_unit], 1e-6)
        name = re.sub(r"\(.*", "", r[k_name])
        name = re.sub(r"^void ", "", name).replace("l3ac::", "")
        n, t = per.get(name, (0, 0.0))
        per[name] = (n + 1, t + ms)
        total += ms
    print(f"# {title}\n\nCommand (under gpurun, 1xB200): `{command}`\n")
    print("Serialised, cold-cache per-launch times: compare SHARES with `op_ms` of the bench line (CUDA events, warm), not absolutes.\n")
    print(f"{sum(n for n, _ in per.values())} launches, {total:.2f} ms in total.\n")
    print("| kernel | launches | sum (ms) | share |\n|---|---|---|---|")
    for name, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {t:.3f} | {100 * t / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
