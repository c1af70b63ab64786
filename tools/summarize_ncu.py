#!/usr/bin/env python
"""Summarise gpurun_out/{launches.csv,prof.ncu-rep,bench*.json} into profiles/<tag>_*.  Run in the build container."""
import csv, json, subprocess, sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
bench_name = sys.argv[2] if len(sys.argv) > 2 else "bench1.json"
rep_name = sys.argv[3] if len(sys.argv) > 3 else "prof.ncu-rep"
src = ROOT / "gpurun_out"
dst = ROOT / "profiles"
dst.mkdir(exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]

out = []
lp = src / "launches.csv"
if lp.exists():
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("shafa::", "")
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        unit = r[hdr.index("Metric Unit")]
        v_ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v_ms
    tot = sum(a[1] for a in agg.values())
    out.append(f"## Launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)\n")
    out.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |")
    out.append("")
    (dst / f"{tag}_launches.csv").write_bytes(lp.read_bytes())

rep = src / rep_name
traffic = {}

def bench_key(kernel: str) -> str:
    """ncu kernel name -> the key bench.py uses for that launch"""
    k = kernel.replace("shafa::", "")
    if k.startswith("k_rle_encode<0, 0>"):
        return "rle_encode_sizes"
    return k.split("<")[0].replace("k_", "", 1)

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out.append("## ncu --set full (one launch per kernel, after warm-up)\n")
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        if name in seen:
            continue
        seen.add(name)
        out.append(f"### {name}\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"| {k} | {r[i]} | {units[i]} |")
        try:
            rd = float(r[hdr.index('dram__bytes_read.sum')]); wr = float(r[hdr.index('dram__bytes_write.sum')])
            u = units[hdr.index('dram__bytes_read.sum')]
            out.append(f"| dram traffic (read+write) | {rd + wr:.4f} | {u} |")
            traffic[bench_key(name)] = int((rd + wr) * UNIT.get(u, 1))
        except Exception:
            pass
        out.append("")
    (dst / f"{tag}_ncu_raw.csv").write_text(raw)
    # per-launch DRAM bytes of each hot kernel on the bench workload: bench.py reports the dominant kernel's as roofline.traffic
    (dst / f"{tag}_traffic.json").write_text(json.dumps({"source": f"profiles/{tag}_ncu_raw.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)",
                                                         "workload": "C2, 1 GiB per GPU", "bytes_per_launch": traffic}, indent=1) + "\n")

for name in (bench_name,):
    p = src / name
    if p.exists() and p.stat().st_size:
        (dst / f"{tag}_bench.json").write_bytes(p.read_bytes())
        j = json.loads(p.read_text().strip().splitlines()[-1])
        out.insert(0, f"## bench.py (not under a profiler)\n\nvalue {j['value']} {j['unit']} round trip, encode {j['encode_gbs']} GB/s, decode {j['decode_gbs']} GB/s, "
                      f"e2e {j['e2e']['value']} GB/s, dominant kernel {j['roofline']['kernel']} at {j['roofline']['achieved']} GB/s "
                      f"= {j['roofline']['frac']} of measured HBM peak {j['roofline']['peak']} GB/s; cpu_baseline {j.get('cpu_baseline', {}).get('value')} GB/s\n")
        out.insert(1, "| kernel | ms/step | algorithmic bytes | GB/s | frac of HBM |\n|---|---|---|---|---|")
        ins = 2
        for k, v in sorted(j["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
            out.insert(ins, f"| {k} | {v['ms_per_step']} | {v.get('algorithmic_bytes', '')} | {v.get('gbs', '')} | {v.get('frac_of_hbm', '')} |")
            ins += 1
        out.insert(ins, "")
(dst / f"{tag}_summary.md").write_text(f"# Profile summary {tag}\n\n" + "\n".join(out) + "\n")
print((dst / f"{tag}_summary.md").read_text()[:3000])
