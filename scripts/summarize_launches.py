"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into a per-kernel table (markdown).
usage: python scripts/summarize_launches.py gpurun_out/launches_X.csv [steps_in_capture] > profiles/X.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit.startswith("ms") else v)   # -> us
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    print("source: `%s` (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and "
          "serialised: read SHARES)\n" % path)
    print("%d launches over %d step(s); sum of kernel time %.3f ms per step\n" % (n, steps, tot / 1e3 / steps))
    print("| kernel | launches/step | ms/step | share |")
    print("|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %.1f | %.3f | %.1f%% |" % (k, c / steps, t / 1e3 / steps, 100 * t / tot))


if __name__ == "__main__":
    main()
