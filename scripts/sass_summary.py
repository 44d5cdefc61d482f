"""Opcode histogram of the tcgen05 / TMA / TMEM instructions per kernel of libfcn8s_sm100.so (cuobjdump -sass), the
tracked proof that the hot kernels are UTCHMMA (tcgen05.mma) / UTMALDG (TMA loads) / LDTM (tcgen05.ld) code.

    python scripts/sass_summary.py [out.txt]          (runs here, no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fcn8s_tensorflow_b200", "libfcn8s_sm100.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCCP", "SYNCS", "ELECT", "HMMA",
         "FFMA", "MUFU", "ATOM", "RED", "STG", "LDG", "STS", "LDS", "SHFL"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for wname in WATCH:
                if op.startswith(wname):
                    kernels[cur][wname] += 1
                    if wname in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
                        kernels[cur][op] += 1
    names = subprocess.run(["c++filt"] + list(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    for k, n in zip(kernels, names):
        demangle[k] = re.sub(r"\(.*", "", n).replace("void ", "")
    tot = collections.Counter()
    lines = ["# SASS opcode histogram of fcn8s_tensorflow_b200/libfcn8s_sm100.so (cuobjdump -sass, sm_100a)", "",
             "UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld,",
             "UTCBAR = tcgen05.commit, UTCCP = tcgen05.cp, UTMASTG = TMA store, SYNCS = mbarrier ops.", ""]
    for k, c in kernels.items():
        if not any(c[w] for w in ("UTCHMMA", "UTMALDG", "LDTM")):
            continue
        detail = ", ".join("%s %d" % (op, n) for op, n in sorted(c.items())
                           if "." in op and op.split(".")[0] in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"))
        lines.append("%s\n    instructions %d | UTCHMMA %d | UTMALDG %d | LDTM %d | UTCBAR %d | UTCCP %d | UTMASTG %d | SYNCS %d | ELECT %d"
                     % (demangle[k], c["_total"], c["UTCHMMA"], c["UTMALDG"], c["LDTM"], c["UTCBAR"], c["UTCCP"],
                        c["UTMASTG"], c["SYNCS"], c["ELECT"]))
        if detail:
            lines.append("    " + detail)
        tot.update({w: c[w] for w in WATCH})
    lines += ["", "library totals over the tensor-core kernels: " + ", ".join("%s %d" % (w, tot[w]) for w in WATCH if tot[w]),
              "", "CUDA-core / elementwise kernels (no tcgen05): " +
              ", ".join(sorted(demangle[k] for k, c in kernels.items() if not any(c[w] for w in ("UTCHMMA", "UTMALDG", "LDTM"))))]
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
