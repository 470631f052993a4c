"""Parse an ncu CSV of one bench step (metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum) into profiles/r2_traffic.json:
per kernel name the launch count, total DRAM bytes and bytes per launch -- the `roofline.traffic` figure of bench.py.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic.csv \\
        python bench.py --steps 1 --warmup 3 --no-other --no-cpu-baseline          (launch-skip / count select one step, see tools/r2_traffic.sh)
    python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/r2_traffic.json --batch 32 --stages seg,depth,warp --depth leres
"""
import argparse
import csv
import gzip
import json
import re
from collections import OrderedDict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def short(name):
    m = re.search(r"(k_[A-Za-z0-9_]+)", name)
    return m.group(1) if m else name.split("(")[0][-60:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv"); ap.add_argument("out")
    ap.add_argument("--batch", type=int, default=32); ap.add_argument("--stages", default="seg,depth,warp"); ap.add_argument("--depth", default="leres")
    ap.add_argument("--steps", type=int, default=1, help="bench steps covered by the capture (totals are divided by it)")
    ap.add_argument("--command", default="", help="the ncu command line, recorded in the output")
    a = ap.parse_args()
    fh = gzip.open(a.csv, "rt", errors="replace") if a.csv.endswith(".gz") else open(a.csv, errors="replace")
    rows = [r for r in csv.reader(fh) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        try:
            k, metric, unit, val = short(r[ix["Kernel Name"]]), r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
        except (ValueError, KeyError, IndexError):
            continue
        d = per.setdefault(k, {"ids": set(), "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "time_us": 0.0})
        d["ids"].add(r[ix["ID"]])
        if metric == "dram__bytes_read.sum":
            d["dram_read_bytes"] += val * UNIT.get(unit, 1.0)
        elif metric == "dram__bytes_write.sum":
            d["dram_write_bytes"] += val * UNIT.get(unit, 1.0)
        elif metric == "gpu__time_duration.sum":
            d["time_us"] += val * UNIT.get(unit, 1.0)
    out = OrderedDict()
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["time_us"]):
        n = len(d["ids"])
        out[k] = {"launches": n / a.steps, "dram_read_bytes": d["dram_read_bytes"] / a.steps, "dram_write_bytes": d["dram_write_bytes"] / a.steps,
                  "dram_bytes_per_launch": (d["dram_read_bytes"] + d["dram_write_bytes"]) / max(1, n), "ncu_time_ms": d["time_us"] * 1e-3 / a.steps}
    res = {"what": "ncu dram__bytes_read.sum + dram__bytes_write.sum per kernel, per bench step (cold-cache, serialised replays; capture = the first "
                   f"{a.steps} step(s) of `bench.py --no-other --no-cpu-baseline`, ncu -c limits the launch count)", "command": a.command, "batch": a.batch,
           "stages": a.stages, "depth": a.depth, "kernels": out}
    json.dump(res, open(a.out, "w"), indent=1)
    for k, v in list(out.items())[:12]:
        print(f"{k:26s} x{v['launches']:6.0f}  {v['dram_bytes_per_launch'] / 1e6:9.1f} MB/launch  {v['ncu_time_ms']:8.2f} ms")


if __name__ == "__main__":
    main()
