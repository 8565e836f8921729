#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libtempme_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st: TMEM), UTMALDG (TMA tensor loads, incl. tile::gather4), UBLKCP (1-D bulk TMA),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus HMMA / FFMA for contrast.  usage: tools/sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "tempme_b200", "csrc", "libtempme_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "ATOM", "RED"]
fn, counts, total = None, collections.OrderedDict(), collections.Counter()
arch = set(re.findall(r"arch = (sm_\w+)", out))
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        total[fn] += 1
        for k in keys:
            if op == k or op.startswith(k + "."):
                counts[fn][k] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(lib)}: SASS mnemonic counts per kernel (cuobjdump -sass); arch {sorted(arch)}")
print(f"{'kernel':78s} {'instrs':>7s} " + " ".join(f"{k:>7s}" for k in keys))
for (f, c), name in zip(counts.items(), demangle):
    name = re.sub(r"\(.*", "", name).replace("tmb::", "")
    print(f"{name[:78]:78s} {total[f]:7d} " + " ".join(f"{c[k]:7d}" for k in keys))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'TOTAL':78s} {sum(total.values()):7d} " + " ".join(f"{tot[k]:7d}" for k in keys))
