#!/usr/bin/env python
"""Turn gpurun_out/{launches_<tag>.csv, prof_<tag>.ncu-rep} into the committed summaries under profiles/:
  profiles/<tag>_launches.md   per-kernel launch count, total and share of the step (ncu gpu__time_duration)
  profiles/<tag>_kernels.md    ncu --set full: time, DRAM bytes, throughput %, occupancy, registers, top stalls
  profiles/<tag>_traffic.json  dram__bytes_read+write per launch for the roofline `traffic` field of bench.py
usage: python tools/summarize_profile.py <tag>"""
import csv, io, json, os, re, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def short(name):
    m = re.search(r"(k_\w+)(<[^>]*>)?", name)
    if not m: return name[:40]
    s = m.group(1)
    t = re.search(r"<(?:slpr::)?(\w+)>", name)
    return s + (f"<{t.group(1)}>" if t else "")

# ---- launch list
rows = [r for r in csv.reader(l for l in open(os.path.join(G, f"launches_{tag}.csv")) if l.startswith('"'))]
h = rows[0]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v if r[ui] in ("us", "usecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v / 1e3
    a = agg.setdefault(short(r[ki]), [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{tag}_launches.md"), "w") as f:
    f.write(f"# ncu launch list, `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` ({tag})\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none`: cold-cache, serialised — compare shares, not absolutes. "
            "Covers the sizing pre-pass, 3 warm-up + 2 timed graph replays, the e2e loops and the instrumented direct-launch frames.\n\n"
            "| kernel | launches | total us | share | us / launch |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {t:.1f} | {100*t/tot:.1f}% | {t/n:.1f} |\n")
    f.write(f"\nraw: `profiles/{tag}_launches.csv`\n")
os.replace(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{tag}_launches.csv")) if False else None
import shutil; shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{tag}_launches.csv"))

# ---- full captures: the default frame (segmented sort) and, when present, the radix sort kernels
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "regs"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst")]
traffic = {}

def process(rep, f):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw))); hh, uu = rr[0], rr[1]
    def col(r, name, default=""):
        return r[hh.index(name)] if name in hh else default
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sr = list(csv.reader(io.StringIO(src)))
    starts = [i for i, r in enumerate(sr) if r and r[0] == "Address"]
    stalls = []
    for si, st in enumerate(starts):
        hdr = sr[st]; end = starts[si + 1] - 1 if si + 1 < len(starts) else len(sr)
        reasons = [(i, x) for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
        acc = {}
        for r in sr[st + 1:end]:
            for i, x in reasons:
                if i < len(r) and r[i].isdigit(): acc[x[6:]] = acc.get(x[6:], 0) + int(r[i])
        t = sum(acc.values()) or 1
        nm = short(sr[st - 1][1]) if st > 0 and len(sr[st - 1]) > 1 else "?"
        stalls.append((nm, ", ".join(f"{k} {100*v/t:.0f}%" for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:4])))
    for n, r in enumerate(rr[2:]):
        name = short(col(r, "Kernel Name"))
        cells = []
        for m, _ in want:
            v = col(r, m); u = uu[hh.index(m)] if m in hh else ""
            try: cells.append(f"{float(v):.2f} {u}".strip())
            except ValueError: cells.append(v)
        st_txt = ""
        for j, (nm, txt) in enumerate(stalls):
            if nm == name:
                st_txt = txt; stalls.pop(j); break
        f.write(f"| `{name}` | " + " | ".join(cells) + f" | {st_txt} |\n")
        def tobytes(m):
            v = float(col(r, m, "0") or 0); u = uu[hh.index(m)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        traffic.setdefault(name, []).append(tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"))

head = "| kernel | " + " | ".join(w[1] for w in want) + " | top stalls |\n|---|" + "---:|" * len(want) + "---|\n"
with open(os.path.join(P, f"{tag}_kernels.md"), "w") as f:
    f.write(f"# ncu --set full, one frame of synth_1m_4k ({tag})\n\n`ncu --set full --clock-control none --import-source on` on `tools/prof_frame.py synth_1m_4k 2` "
            "(second frame, direct launches, default = segmented sort). Times under ncu are serialised and slower than in situ.\n\n" + head)
    process(os.path.join(G, f"prof_{tag}.ncu-rep"), f)
    rep2 = os.path.join(G, f"prof_{tag}_radix.ncu-rep")
    if os.path.exists(rep2):
        f.write("\n## radix sort kernels (`SLPR_FLAG_RADIX_SORT`, the general path: paths of any length)\n\n" + head)
        process(rep2, f)
json.dump({"workload": "synth_1m_4k", "source": f"profiles/{tag}_kernels.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
           "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in traffic.items()}}, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_launches.md")).read()); print(open(os.path.join(P, f"{tag}_kernels.md")).read())
