#!/usr/bin/env python
"""profiles/<name>.json from an ncu metrics CSV of `bench.py` (single-pass metrics: no kernel replay), e.g.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum \\
      --clock-control none -k regex:'k_' --csv --log-file launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras
  python tools/ncu_traffic.py launches.csv profiles/r2_traffic_10gb.json --chunk 1048576 --level 3 --nchunks 9481

bench.py reads the result for `roofline.traffic` (DRAM bytes per launch of the dominant kernel) when its own
configuration matches the one recorded here.  Per kernel name: the launch with the longest duration."""
import argparse
import csv
import json
import re
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("out")
    ap.add_argument("--chunk", type=int, default=1 << 20)
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--nchunks", type=int, required=True)
    ap.add_argument("--how", default="ncu single-pass metrics over bench.py --steps 1 --warmup 1 (kernels serialised by ncu)")
    a = ap.parse_args()
    rows = [r for r in csv.reader(open(a.csv, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    per = {}
    for r in rows:
        lid, name, metric, value = r[0], r[4], r[-3], r[-1]
        m = re.match(r"(?:void )?([A-Za-z0-9_]+(?:<[^>]*>)?)", name)
        key = m.group(1).replace("(int)", "").replace("(bool)", "") if m else name
        d = per.setdefault((key, lid), {})
        try:
            v = float(value.replace(",", ""))
            if v == v:                       # (a launch ncu was cut off in reports nan)
                d[metric] = v
        except ValueError:
            pass
    best = {}
    for (key, lid), d in per.items():
        if "gpu__time_duration.sum" not in d:
            continue
        if key not in best or d["gpu__time_duration.sum"] > best[key]["gpu__time_duration.sum"]:
            best[key] = d
    out = {"config": {"chunk_bytes": a.chunk, "level": a.level, "nchunks": a.nchunks}, "how": a.how, "kernels": {}}
    for key, d in sorted(best.items()):
        out["kernels"][key] = {"duration_ms": round(d["gpu__time_duration.sum"] / 1e6, 3),
                               "dram_bytes_read": int(d.get("dram__bytes_read.sum", 0)), "dram_bytes_write": int(d.get("dram__bytes_write.sum", 0)),
                               "l2_hit_pct": round(d.get("lts__t_sector_hit_rate.pct", 0), 2), "warp_instructions": int(d.get("smsp__inst_executed.sum", 0))}
    json.dump(out, open(a.out, "w"), indent=1)
    for k, v in out["kernels"].items():
        print("%-28s %9.2f ms  dram %8.2f GB read %8.2f GB written  L2 hit %5.1f %%" % (k, v["duration_ms"], v["dram_bytes_read"] / 1e9, v["dram_bytes_write"] / 1e9, v["l2_hit_pct"]))


if __name__ == "__main__":
    sys.exit(main())
