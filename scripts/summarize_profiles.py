#!/usr/bin/env python
"""Turn the csv pages the GPU pass brings back (gpurun_out/) into the summaries committed under profiles/.
    python scripts/summarize_profiles.py r02"""
import collections
import csv
import json
import os
import sys

csv.field_size_limit(10 ** 9)
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launch_list(w):
    src = os.path.join(G, f"{R}_launches_{w}_bench.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                agg[d["Kernel Name"].split("(")[0].replace("void ", "").replace("mb::", "")[:70]].append(float(d["Metric Value"]) / 1000.0)
            except ValueError:
                pass
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, f"{R}_launches_{w}_bench.md"), "w") as f:
        f.write(f"# {R} — ncu launch list, workload {w.upper()} (`MOLLYB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none "
                f"-c 900 python bench.py --workload {w} --steps 40 --warmup 4 --no-e2e --no-cpu-baseline`)\n\n"
                "Stream mode: the gated rebuild kernels are enqueued every step and return at once unless a rebuild is due, so their "
                "means mix a few real runs with many ~2-3 us no-ops (max = a real run). Times are ncu's serialised, cold-cache "
                "per-launch durations: use the SHARES, the absolute step time is bench.py's.\n\n"
                "| kernel | launches | mean us | max us | total ms | share |\n|---|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
            f.write(f"| {k} | {len(v)} | {sum(v) / len(v):.2f} | {max(v):.1f} | {sum(v) / 1000:.3f} | {100 * sum(v) / tot:.1f}% |\n")
    import shutil
    shutil.copy(src, os.path.join(P, f"{R}_launches_{w}_bench.csv"))


def raw(name):
    src = os.path.join(G, f"{R}_{name}.raw.csv")
    if not os.path.exists(src):
        return None
    rows = list(csv.reader(open(src)))
    return dict(zip(rows[0], rows[2]))


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_local_ld.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def table(d):
    out = "| metric | value |\n|---|---:|\n"
    for k in KEYS:
        if k in d and d[k] not in (None, ""):
            out += f"| `{k}` | {d[k]} |\n"
    st, tot = {}, 0
    for k, v in d.items():
        if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
            st[k[33:]] = int(v)
            tot += int(v)
    if tot:
        out += "\nWarp-state samples: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]) + ".\n"
    return out


def source_summary():
    src = os.path.join(G, f"{R}_force_c2.source.csv")
    if not os.path.exists(src):
        return ""
    rows = list(csv.reader(open(src)))[2:]
    op = collections.Counter()
    cls = collections.Counter()
    for r in rows:
        t = r[1].strip().split()
        if not t:
            continue
        o = t[1] if t[0].startswith("@") else t[0]
        op[o.split(".")[0] + ("." + o.split(".")[1] if o.startswith(("LDS", "LDG", "UBLKCP", "SYNCS", "ATOMS", "REDUX")) and "." in o else "")] += int(r[5])
        cls[int(r[5])] += int(r[5])
    tot = sum(op.values())
    s = f"\nSASS opcode mix (warp instructions executed, source page, {tot / 1e6:.1f} M): " + ", ".join(
        f"`{k}` {v / 1e6:.2f} M" for k, v in op.most_common(28)) + ".\n"
    top = sorted(range(len(rows)), key=lambda i: -int(rows[i][4]))[:12]
    s += "\nMost-sampled instructions (warp-state samples; the sample sits on the instruction that WAITS):\n\n| SASS | samples | executed |\n|---|---:|---:|\n"
    for i in top:
        s += f"| `{rows[i][1].strip()[:80]}` | {rows[i][4]} | {rows[i][5]} |\n"
    return s


if __name__ == "__main__":
    for w in ("c2", "c3"):
        launch_list(w)
    d = raw("force_c2")
    if d:
        open(os.path.join(P, f"{R}_force_kernel_c2.ncu.md"), "w").write(
            f"# {R} — `ncu --set full` of brick_force_kernel<float, COUL_NONE, UNIFORM, plain cutoff, no energy>, workload C2\n\n"
            f"Command: `MOLLYB200_NO_GRAPH=1 ncu --set full --import-source on --clock-control none -k regex:brick_force_kernel -s 5 -c 1 "
            f"python bench.py --workload c2 --steps 12 --warmup 4 --no-e2e --no-cpu-baseline`; pages `--page raw --csv` / `--page source --csv`.\n\n"
            + table(d) + source_summary())
        json.dump({"dram_bytes_per_launch": int(float(d["dram__bytes_read.sum"]) * 1e6 + float(d["dram__bytes_write.sum"]) * 1e6),
                   "source": f"profiles/{R}_force_kernel_c2.ncu.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"},
                  open(os.path.join(P, "force_kernel_c2.json"), "w"))
    for name, title in (("build_c2", "build_lists_kernel<float, lists, no exclusions> (one real rebuild)"), ("k1_c2", "vv_kick_drift_kernel<float>")):
        d = raw(name)
        if d:
            open(os.path.join(P, f"{R}_{name}.ncu.md"), "w").write(f"# {R} — `ncu --set full` of {title}, workload C2\n\n" + table(d))
    for f in (f"{R}_bench_c2.json", f"{R}_bench_c2_reference_arm.json", f"{R}_gputests.log"):
        if os.path.exists(os.path.join(G, f)):
            import shutil
            shutil.copy(os.path.join(G, f), os.path.join(P, f))
