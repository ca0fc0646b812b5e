"""Turns the scratch captures of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <tag> <round-name>

Writes profiles/<round>_launches.md (per-kernel share of a short bench run, from the ncu launch list),
profiles/<round>_scan_ncu.md / _colstats_ncu.md (selected raw metrics of the --set full captures) and
profiles/scan_traffic.json (dram bytes per scan launch, read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

RAW_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def launches(tag, rnd):
    path = os.path.join(OUT, "launches_%s.csv" % tag)
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if r[0] == "ID")
    body = rows[rows.index(hdr) + 1:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot, cnt = collections.Counter(), collections.Counter()
    for r in body:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = r[ix["Kernel Name"]].split("(")[0]
        v = to_float(r[ix["Metric Value"]])
        u = r[ix["Metric Unit"]]
        v = v / 1e3 if u.startswith("n") else v * 1e3 if u.startswith("m") else v
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    lines = ["# %s: kernel launch list of `BMG_COLSTATS_SERVER=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline`" % rnd, "",
             "(launch-per-move form of the column statistics: ncu serialises kernels, which the default persistent",
             "`k_colstats_server` cannot run under; the work items are the same.)", "",
             "Captured with `ncu --metrics gpu__time_duration.sum --clock-control none` (tools/profile_round.sh);",
             "per-launch times under ncu are serialised and cold-cache, so only the SHARES are meaningful.", "",
             "%d launches, %.1f us of kernel time in total." % (sum(cnt.values()), total), "",
             "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, v in tot.most_common():
        lines.append("| `%s` | %d | %.1f | %.1f%% | %.2f |" % (k, cnt[k], v, 100 * v / total, v / cnt[k]))
    open(os.path.join(PROF, "%s_launches.md" % rnd), "w").write("\n".join(lines) + "\n")
    return tot, cnt


def raw(name, title, rnd, out_name):
    path = os.path.join(OUT, name)
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    hdr, units, r = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = ["# %s: %s" % (rnd, title), "", "`ncu --set full --clock-control none --import-source on`, one launch after warm-up; raw page.",
             "", "| metric | value | unit |", "|---|---:|---|"]
    for m in RAW_METRICS:
        if m in ix:
            lines.append("| %s | %s | %s |" % (m, r[ix[m]], units[ix[m]]))
    lines += ["", "Warp stall reasons (warps per issue-active cycle, > 0.05):", "", "| reason | ratio |", "|---|---:|"]
    for i, h in enumerate(hdr):
        if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            v = to_float(r[i])
            if v is not None and v > 0.05:
                lines.append("| %s | %.3f |" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    open(os.path.join(PROF, out_name), "w").write("\n".join(lines) + "\n")
    rd, wr = to_float(r[ix["dram__bytes_read.sum"]]), to_float(r[ix["dram__bytes_write.sum"]])
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return rd * mult[units[ix["dram__bytes_read.sum"]]] + wr * mult[units[ix["dram__bytes_write.sum"]]]


def main():
    tag, rnd = sys.argv[1], sys.argv[2]
    os.makedirs(PROF, exist_ok=True)
    launches(tag, rnd)
    traffic = {}
    t = raw("scan_%s_raw.csv" % tag, "k_scan_dots_imma at C2 (n=5,000 x m=100,000; 125.84 MB of packed genotypes per launch)", rnd,
            "%s_scan_ncu.md" % rnd)
    if t:
        traffic["C2"] = t
    t = raw("scan1m_%s_raw.csv" % tag, "k_scan_dots_imma at n=5,000 x m=1,000,000 (1.2584 GB per launch, 10x the L2)", rnd,
            "%s_scan1m_ncu.md" % rnd)
    if t:
        traffic["C2x"] = t
    raw("colstats_%s_raw.csv" % tag, "k_column_stats_inline (one proposal's column statistics; latency-bound)", rnd,
        "%s_colstats_ncu.md" % rnd)
    raw("scan2_%s_raw.csv" % tag, "k_scan_dots_imma2 (two residuals per pass; shard groups) at n=50,000 x m=200,000 (2.5 GB of packed genotypes per launch)",
        rnd, "%s_scan2_ncu.md" % rnd)
    if traffic:
        json.dump(traffic, open(os.path.join(PROF, "scan_traffic.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()
