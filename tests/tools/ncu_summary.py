"""Summarise ncu captures into profiles/ (tracked): per-kernel time shares from a launch list
(`--metrics gpu__time_duration.sum`) and key counters from a `--set full` report.

  python tests/tools/ncu_summary.py <tag>      # reads gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_prof.ncu-rep
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_shared_atom.sum",
        "smsp__inst_executed_op_global_red.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(tag, out):
    p = ROOT / "gpurun_out" / f"{tag}_launches.csv"
    if not p.exists():
        return
    lines = [l for l in open(p) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        k = row["Kernel Name"].split("(")[0][:80]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out.append(f"## Launch list `{p.name}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
    out.append("| kernel | launches | total ms | us / launch | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.002:
            continue
        out.append(f"| `{k}` | {v[0]} | {v[1]:.3f} | {v[1] / v[0] * 1e3:.1f} | {v[1] / tot:.1%} |")
    out.append(f"\ntotal {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")


def full(tag, out):
    p = ROOT / "gpurun_out" / f"{tag}_prof.ncu-rep"
    if not p.exists():
        return
    raw = subprocess.run(["ncu", "-i", str(p), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    out.append(f"## Full capture `{p.name}` (ncu --set full --clock-control none --import-source on)\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        out.append(f"### `{name}`\n\n| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in idx:
                out.append(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
        # which execution pipe the issue slots go to (fma / alu / xu / lsu ...)
        for k in hdr:
            if k not in KEYS and ("inst_executed_pipe_" in k or "_cycles_active" in k and "pipe_" in k) and \
                    "pct_of_peak_sustained_active" in k and "tensor" not in k:
                try:
                    if float(r[idx[k]].replace(",", "")) >= 1.0:
                        out.append(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
                except ValueError:
                    pass
        out.append("")


STAGE_OF = {"project_fwd_kernel": "project_fwd", "tile_scan_kernel": "tile_scan", "scatter_kernel": "scatter",
            "tile_sort_kernel": "tile_sort", "tile_sort_warp_kernel": "tile_sort", "blend_fwd_kernel": "blend_fwd", "blend_fwd_tma_kernel": "blend_fwd",
            "blend_bwd_kernel": "blend_bwd", "blend_fwd_warp_kernel": "blend_fwd", "blend_bwd_warp_kernel": "blend_bwd", "project_bwd_kernel": "project_bwd",
            "blend_fwd_pair_kernel": "blend_fwd", "blend_bwd_tall_kernel": "blend_bwd", "blend_bwd_pair_kernel": "blend_bwd",
            "acc_clear_kernel": "acc_clear"}


def traffic(tag):
    """profiles/traffic.json: measured DRAM bytes (read + write) per launch of each stage, for bench.py's roofline.traffic."""
    import json
    p = ROOT / "gpurun_out" / f"{tag}_prof.ncu-rep"
    if not p.exists():
        return
    raw = subprocess.run(["ncu", "-i", str(p), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].split("<")[0].split("::")[-1].split()[-1]
        st = STAGE_OF.get(name)
        if st is None:
            continue
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[idx[k]].replace(",", "")) * mult.get(units[idx[k]], 1.0)
        acc.setdefault(st, []).append(b)
    out = {k: sum(v) / len(v) for k, v in acc.items()}
    out["_source"] = f"{p.name}: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean over captured launches"
    (ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=1))


def main():
    tag = sys.argv[1]
    out = [f"# ncu summary {tag}\n", "Workload: `python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline` "
           "(c2: 1.0 M surfels, 1920x1080) on one B200.  Numbers under the profiler are never bench values; "
           "compare shares.\n"]
    launches(tag, out)
    full(tag, out)
    dst = ROOT / "profiles" / f"{tag}_ncu_summary.md"
    dst.write_text("\n".join(out) + "\n")
    traffic(tag)
    print("wrote", dst)


if __name__ == "__main__":
    main()
