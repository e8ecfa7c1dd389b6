#!/usr/bin/env python
"""Disassemble libvz_b200.so and count, per kernel, the SASS mnemonics that prove which hardware path it
uses: UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (tcgen05.ld / st), UTCBAR
(tcgen05.commit), HMMA / IMMA (legacy mma.sync), IDP (IDP.4A byte dot product).  Writes profiles/sass_summary.txt.

    python tools/sass_summary.py            # needs cuobjdump (CUDA toolkit); no GPU
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "vision-zephyr_b200", "libvz_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA", "IMMA", "IDP", "MUFU.EX2"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], check=True, capture_output=True, text=True).stdout
    fn, counts, sizes = None, collections.OrderedDict(), {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            fn = re.sub(r"\(anonymous namespace\)::", "", fn)
            counts[fn] = collections.Counter()
            sizes[fn] = 0
            continue
        if fn is None or "/*" not in line:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        sizes[fn] += 1
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + ".") or (o == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                counts[fn][o] += 1
    out = ["# SASS summary of vision-zephyr_b200/libvz_b200.so (cuobjdump -sass, sm_100a), per kernel",
           "# columns: instructions | " + " ".join(OPS), ""]
    tot = collections.Counter()
    for fn, c in counts.items():
        tot.update(c)
        short = fn if len(fn) < 110 else fn[:107] + "..."
        out.append(f"{sizes[fn]:7d} | " + " ".join(f"{o}={c[o]}" for o in OPS if c[o]) + f"  :: {short}")
    out.append("")
    out.append("TOTAL " + " ".join(f"{o}={tot[o]}" for o in OPS))
    legacy = [fn for fn, c in counts.items() if c["HMMA"] or c["IMMA"]]
    out.append("kernels with legacy mma.sync (HMMA / IMMA): " + (", ".join(f.split("(")[0] for f in legacy) or "none"))
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out[-3:]))
    print("written", path)


if __name__ == "__main__":
    sys.exit(main())
