"""Per-kernel table of ONE train step from an ncu launch list.

    FCN8_GRAPHS=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file X.csv \
        python bench.py --profile --precision P --steps 1 --warmup 1
    python scripts/step_table.py X.csv P profiles/rNN_launches_P_TAG.md [profiles/rNN_step_kernels_TAG.json]

The last complete step of the capture is used (a step starts at set_step_scalars_kernel).  The JSON (appended to /
updated per precision) is what bench.py reads for roofline.traffic.
"""
import collections
import csv
import json
import os
import re
import sys


def main():
    path, precision, out_md = sys.argv[1:4]
    out_json = sys.argv[4] if len(sys.argv) > 4 else None
    lines = [l for l in open(path) if not l.startswith("==")]
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = launches.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
                                               .replace("fcn8::", "")})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if unit == "ns" else (v * 1e3 if unit.startswith("ms") else v)
        elif m.startswith("dram__bytes"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d["dram"] = d.get("dram", 0.0) + v * scale
        elif m.startswith("sm__pipe_tensor"):
            d["tensor"] = v
    seq = list(launches.values())
    starts = [i for i, d in enumerate(seq) if d["name"].startswith("set_step_scalars")]
    step = seq[starts[-2]:starts[-1]]
    agg = collections.OrderedDict()
    for d in step:
        a = agg.setdefault(d["name"], {"launches": 0, "us": 0.0, "dram": 0.0, "tw": 0.0})
        a["launches"] += 1
        a["us"] += d.get("us", 0.0)
        a["dram"] += d.get("dram", 0.0)
        a["tw"] += d.get("tensor", 0.0) * d.get("us", 0.0)
    tot = sum(a["us"] for a in agg.values())
    with open(out_md, "w") as f:
        f.write("source: `%s` (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none; eager mode "
                "FCN8_GRAPHS=0, last complete step of the capture; per-launch times are cold-cache and serialised: "
                "read SHARES)\n\n" % path)
        f.write("precision %s, c2 = 4 x 512x1024x3, 20 classes: %d launches, sum of kernel time %.3f ms per step\n\n"
                % (precision, len(step), tot / 1e3))
        f.write("| kernel | launches/step | us/step | share | DRAM MB/step | tensor pipe % (time-weighted, of ncu's "
                "nominal peak) |\n|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            f.write("| `%s` | %d | %.1f | %.1f%% | %.1f | %.1f |\n"
                    % (k, a["launches"], a["us"], 100 * a["us"] / tot, a["dram"] / 1e6, a["tw"] / max(a["us"], 1e-9)))
    if out_json:
        data = json.load(open(out_json)) if os.path.exists(out_json) else {}
        data[precision] = {k: {"launches": a["launches"], "us": a["us"], "dram_bytes": a["dram"],
                               "tensor_pct_time_weighted": a["tw"] / max(a["us"], 1e-9)} for k, a in agg.items()}
        json.dump(data, open(out_json, "w"), indent=1)
    print("%s: %d launches, %.3f ms" % (precision, len(step), tot / 1e3))


if __name__ == "__main__":
    main()
