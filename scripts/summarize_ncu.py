#!/usr/bin/env python
"""Turn gpurun_out/<tag>/ (written by scripts/gpu_profile.sh) into committed summaries:

  profiles/<tag>_launches_<workload>.md   per-kernel launch counts, time and SHARE (ncu, cold cache)
  profiles/<tag>_full_<workload>.md       key `ncu --set full` metrics per captured launch
  profiles/traffic.json                   dram bytes per launch per kernel tag (read by bench.py)

Usage: python scripts/summarize_ncu.py <tag>
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# kernel-name regex -> the launch tag bench.py uses (fdm_b200/csrc LaunchScope tags)
TAGS = [
    (r"k_fgh_rhs", "ns_fgh_rhs"), (r"k_fgh", "ns_fgh"), (r"k_rhs", "ns_rhs"), (r"k_update", "ns_update"),
    (r"k_bound_lid", "ns_bound_lid"), (r"k_bound_mirror", "ns_bound_mirror"), (r"k_bound_p", "ns_bound_p"),
    (r"k_cyl_fgh<\(bool\)0>|k_cyl_fgh<0>", "nscyl_fgh"), (r"k_cyl_fgh", "nscyl_lfgh"), (r"k_cyl_rhs", "nscyl_rhs"),
    (r"k_cyl_update", "nscyl_update"), (r"k_cyl_bound_r", "nscyl_bound_r"), (r"k_cyl_bound_z", "nscyl_bound_z"),
    (r"k_cyl_bound_p", "nscyl_bound_p"), (r"k_tridiag_rows", "cyl_r_tridiag"),
]

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dsmem"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smem_pipe_pct"),
])


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
    return v * mult


def to_us(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit.lower(), 1)


def short_name(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    return name.replace("fdmb::", "")


def launches(tag, wl):
    path = os.path.join(ROOT, "gpurun_out", tag, f"launches_{wl}.csv")
    if not os.path.exists(path):
        return
    text = open(path, errors="ignore").read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short_name(r["Kernel Name"])
        us = to_us(r["Metric Value"], r["Metric Unit"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += us
    ours = {k: v for k, v in agg.items() if "k_" in k}
    tot = sum(v[1] for v in ours.values()) or 1.0
    out = [f"# ncu launch list: bench.py --workload {wl} --steps 3 --warmup 3 ({tag})", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and",
           "serialised; compare SHARES with bench.py's live CUDA-event numbers, not absolutes.", "",
           "| kernel (this library) | launches | total us | us/launch | share of library kernels |", "|---|---:|---:|---:|---:|"]
    for k, (n, us) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {us:.1f} | {us / n:.2f} | {100 * us / tot:.1f}% |")
    others = {k: v for k, v in agg.items() if "k_" not in k}
    if others:
        out += ["", "Other launches in the same process (torch RNG/fill for the synthetic inputs, not on the path):", ""]
        for k, (n, us) in sorted(others.items(), key=lambda kv: -kv[1][1])[:6]:
            out.append(f"- `{k[:90]}` x{n}: {us:.1f} us")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches_{wl}.md"), "w").write("\n".join(out) + "\n")


def full(tag, wl, traffic):
    rep = os.path.join(ROOT, "gpurun_out", tag, f"full_{wl}.ncu-rep")
    rawcsv = os.path.join(ROOT, "gpurun_out", tag, f"full_{wl}.raw.csv")
    if os.path.exists(rawcsv):
        raw = open(rawcsv, errors="ignore").read()
    elif os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full: bench.py --workload {wl} ({tag})", "",
           "`ncu --set full --clock-control none --import-source on`; one column per captured launch.", ""]
    cols = []
    for r in rows[2:]:
        d = {}
        name = short_name(r[hdr.index("Kernel Name")])
        d["kernel"] = name
        for m, short in METRICS.items():
            if m in hdr:
                i = hdr.index(m)
                d[short] = (r[i], units[i])
        cols.append(d)
    out.append("| metric | " + " | ".join(f"`{c['kernel'][:48]}`" for c in cols) + " |")
    out.append("|---|" + "---:|" * len(cols))
    for m, short in METRICS.items():
        cells = []
        for c in cols:
            if short not in c:
                cells.append("-"); continue
            v, u = c[short]
            if short in ("dram_rd", "dram_wr", "l2_bytes"):
                cells.append(f"{to_bytes(v, u) / 1e6:.1f} MB")
            elif short == "time":
                cells.append(f"{to_us(v, u):.1f} us")
            else:
                try:
                    cells.append(f"{float(v.replace(',', '')):.2f}")
                except ValueError:
                    cells.append(v)
        out.append(f"| {short} (`{m}`) | " + " | ".join(cells) + " |")
    open(os.path.join(ROOT, "profiles", f"{tag}_full_{wl}.md"), "w").write("\n".join(out) + "\n")
    # traffic per launch, keyed the way bench.py tags kernels
    tr = traffic.setdefault(wl, {})
    seen = {}
    for c in cols:
        if "dram_rd" not in c:
            continue
        b = to_bytes(*c["dram_rd"]) + to_bytes(*c["dram_wr"])
        seen.setdefault(c["kernel"], []).append(b)
    tr["_by_kernel_name"] = {k: sum(v) / len(v) for k, v in seen.items()}
    for k, v in seen.items():
        for rx, t in TAGS:
            if re.search(rx, k):
                tr[t] = sum(v) / len(v)
                break
    if wl.startswith("cube") and len(cols) == 5:
        # the capture window is one solve: x fwd, y fwd, z fwd/divide/inv, y inv, x inv (LaunchScope tags of lapl_cube.cu)
        for c, t in zip(cols, ["cube_x_fwd", "cube_y_fwd", "cube_z_fwd_div_inv", "cube_y_inv", "cube_x_inv"]):
            if "dram_rd" in c:
                tr[t] = to_bytes(*c["dram_rd"]) + to_bytes(*c["dram_wr"])


def main():
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for wl in ("cube127", "cube255", "nscube255", "cube511", "cube1023", "nscyl128", "cyl128"):
        launches(tag, wl)
        full(tag, wl, traffic)
    traffic["_source"] = f"ncu --set full captures of {tag}; bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)"
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
