#!/usr/bin/env python
"""Turn the ncu captures in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py r01

Writes profiles/<round>_launches.csv (per-kernel totals of the launch list), profiles/<round>_ncu_kernels.csv
(one row per --set full capture: duration, DRAM bytes, tensor-pipe %, L2->SM bytes, registers) and
profiles/ncu_traffic.json (dram bytes per launch by kernel class, read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, os.environ.get("PROF_IN", "gpurun_out"))   # PROF_IN: directory the captures were written to
PROF = os.path.join(ROOT, "profiles")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name, val, unit = r[4], float(r[-1]), r[-2]
        us = val / 1000.0 if unit.startswith("n") else val
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_captured_time"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], f"{v[1]:.1f}", f"{v[1] / v[0]:.2f}", f"{v[1] / tot:.4f}"])
    print(f"launch list: {len(rows)} launches, {tot / 1e3:.2f} ms captured")


def full_captures(tag):
    out_rows, traffic = [], {}
    # either the reports themselves or their raw pages exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv >
    # x.raw.csv`: a report embeds the whole cubin, ~16 MB each, and gpurun returns at most 64 MiB)
    names = sorted(f for f in os.listdir(OUT) if f.endswith(".raw.csv") or
                   (f.endswith(".ncu-rep") and not os.path.exists(os.path.join(OUT, f[:-8] + ".raw.csv"))))
    for rep in names:
        if rep.endswith(".raw.csv"):
            text = open(os.path.join(OUT, rep)).read()
            rep = rep[:-8] + ".ncu-rep"
        else:
            text = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"], capture_output=True,
                                  text=True).stdout
        rows = list(csv.reader(text.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}  # -> MB / us
        for r in rows[2:]:
            rec = {"report": rep, "kernel": r[hdr.index("Kernel Name")]}
            for m in METRICS:
                if m not in hdr:
                    rec[m] = ""
                    continue
                v, u = r[hdr.index(m)], units[hdr.index(m)]
                if (m.startswith("dram__bytes") or m.startswith("l1tex__m_xbar") or m == "gpu__time_duration.sum") and u in scale:
                    v = f"{float(v.replace(',', '')) * scale[u]:.3f}"   # normalised: MB and microseconds
                rec[m] = v
            out_rows.append(rec)
    if not out_rows:
        return
    with open(os.path.join(PROF, f"{tag}_ncu_kernels.csv"), "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=["report", "kernel"] + METRICS)
        w.writeheader()
        w.writerows(out_rows)
    # DRAM traffic per launch, keyed by bench configuration (model_method_bN) and bench.py's kernel class
    cfgs = {"c2": "vit_b32_kadaptation_b256", "L197": "vit_b16_lora_b512", "L257": "vit_l14_kadaptation_b256"}
    for rec in out_rows:
        try:
            b = (float(rec["dram__bytes_read.sum"]) + float(rec["dram__bytes_write.sum"])) * 1e6
        except ValueError:
            continue
        k, rep = rec["kernel"], rec["report"]
        if rep.startswith("one_"):   # tools/one_gemm.py captures: one GEMM class per report, C2 shapes
            traffic.setdefault(cfgs["c2"], {})["gemm_" + rep[4:].split(".")[0]] = b
        elif rep.startswith("hr_"):  # tools/attn_bench.py captures of the head-resident kernels: hr_<fwd|bwd>_L<L>
            _, kern, ltag = rep.split(".")[0].split("_")
            traffic.setdefault(cfgs[ltag], {})["attn_" + kern] = b
        elif "attn_fwd_tc" in k:
            traffic.setdefault(cfgs["c2"], {})["attn_fwd"] = b
        elif "attn_bwd_tc" in k:
            traffic.setdefault(cfgs["c2"], {})["attn_bwd"] = b
    with open(os.path.join(PROF, "ncu_traffic.json"), "w") as fh:
        json.dump(traffic, fh, indent=1, sort_keys=True)
    print(f"full captures: {len(out_rows)} kernels")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    full_captures(tag)
