#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMA / TMEM path (B200_PROFILING.md), from
`cuobjdump -sass` of the built libb2v.so (a build artefact, git-ignored -- this listing is the tracked evidence).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video-to-video-diffusion_b200", "libb2v.so")
PATTERNS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOM", "UCGABAR", "SYNCS",
            "HMMA", "ATOMG", "REDG", "ATOMS", "LDGSTS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for p in PATTERNS:
            if op.startswith(p):
                kernels[cur][p] += 1
                break
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# cuobjdump -sass libb2v.so (sm_100a), built from {sha}; counts of instructions per kernel")
    print(f"# columns: {' '.join(PATTERNS)}  (UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA load/store, LDTM = tcgen05.ld,")
    print("#          UTCBAR = tcgen05.commit, UCGABAR = cluster barrier, SYNCS = mbarrier, HMMA = legacy mma.sync)")
    for (name, c), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm).replace("b2v::", "")
        cols = " ".join(f"{p}={c[p]}" for p in PATTERNS if c[p])
        print(f"{short:<48s} insts={c['_total']:<6d} {cols}")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("# total: " + " ".join(f"{p}={tot[p]}" for p in PATTERNS))


if __name__ == "__main__":
    sys.exit(main())
