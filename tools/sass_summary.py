#!/usr/bin/env python
"""profiles/<tag>_sass_summary.md: per kernel of libslpr.so — registers / shared memory / spills (cuobjdump
--dump-resource-usage) and counts of the SASS mnemonics that matter for this code base (128-bit global accesses,
fused multiply-adds — the arithmetic policy forbids contraction, so FFMA may only appear inside IEEE division /
square-root sequences and integer-to-float address arithmetic —, atomics, shuffles, bulk copies, mbarrier ops).
usage: python tools/sass_summary.py <tag>"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
so = os.path.join(ROOT, "vkscanlinepr_b200", "libslpr.so")
res = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)))
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = {}
names = list(usage)
d = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
for n, dn in zip(names, d):
    m = re.search(r"(k_\w+)(<[^>(]*>)?", dn)
    dem[n] = (m.group(1) + (m.group(2) or "")) if m else dn[:40]
pat = {"LDG.128": r"\bLDG\.E(\.\w+)*\.128", "STG.128": r"\bSTG\.E(\.\w+)*\.128", "LDG": r"\bLDG\.", "STG": r"\bSTG\.", "FFMA": r"\bFFMA\b", "FADD": r"\bFADD\b",
       "FMUL": r"\bFMUL\b", "MUFU": r"\bMUFU\.", "RED/ATOM": r"\b(RED|ATOMG|ATOM)\.", "SHFL": r"\bSHFL\.", "LDGSTS": r"\bLDGSTS", "UBLKCP": r"\bUBLKCP",
       "SYNCS": r"\bSYNCS\."}
rows = []
for blk in re.split(r"\n\s*Function : ", sass)[1:]:
    name = blk.split("\n", 1)[0].strip()
    if name not in usage:
        continue
    cnt = {k: len(re.findall(p, blk)) for k, p in pat.items()}
    rows.append((dem[name], usage[name], cnt))
rows.sort(key=lambda r: r[0])
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md"), "w") as f:
    f.write(f"# SASS / resource summary of libslpr.so ({tag})\n\n`cuobjdump --dump-resource-usage` and mnemonic counts from `cuobjdump -sass` (static counts). "
            "The arithmetic policy is IEEE fp32 without contraction (`-fmad=false`, `__f*_rn` intrinsics): FFMA appears only inside the IEEE division / "
            "square-root sequences (`__fdiv_rn`, `__fsqrt_rn` expand to MUFU + FFMA refinement) and float index arithmetic — never in a LERP.\n\n"
            "| kernel | regs | stack B | smem B | local B | " + " | ".join(pat) + " |\n|---|---:|---:|---:|---:|" + "---:|" * len(pat) + "\n")
    for n, u, c in rows:
        f.write(f"| `{n}` | {u[0]} | {u[1]} | {u[2]} | {u[3]} | " + " | ".join(str(c[k]) for k in pat) + " |\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md")).read()[:3000])
