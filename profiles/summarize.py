#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1_v4.csv profiles/launches_r1_v4.md [steps]
    python profiles/summarize.py raw gpurun_out/prof_field_r1.ncu-rep profiles/ncu_full_r1.md
"""
import collections
import csv
import subprocess
import sys

RAW_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def launches(src, dst, steps=None):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows:
        v = float(r["Metric Value"].replace(",", "")) / 1e6
        agg[r["Kernel Name"]][0] += 1
        agg[r["Kernel Name"]][1] += v
        tot += v
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("source: `%s`, %d launches, %.3f ms of kernel time" % (src, len(rows), tot))
        if steps:
            f.write(" (%s steps incl. warm-up / capture: compare SHARES, per-launch times are cold-cache)" % steps)
        f.write("\n\n| total ms | launches | share | avg ms | kernel |\n|---:|---:|---:|---:|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write("| %.3f | %d | %.1f%% | %.4f | `%s` |\n" % (t, n, 100 * t / tot, t / n, k[:110].replace("|", "/")))


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full summary of `%s`\n" % src)
        for r in rows[2:]:
            f.write("\n## %s  (launch id %s)\n\n| metric | value | unit |\n|---|---:|---|\n" % (r[idx["Kernel Name"]][:100], r[idx["ID"]]))
            for m in RAW_METRICS:
                if m in idx:
                    f.write("| %s | %s | %s |\n" % (m, r[idx[m]], units[idx[m]]))


def traffic(src, dst, samples_per_launch=None):
    """profiles/ncu_traffic.json: measured DRAM bytes per launch of every kernel in a `ncu --set full` capture (mean over the
    captured launches), keyed by the C-ABI entry point bench.py times.  bench.py puts the entry of its dominant kernel into
    `roofline.traffic`."""
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    entry_of = {"encode_backward_warpagg_kernel": "ngp_grid_scatter_samples", "field_forward_kernel": "ngp_field_forward",
                "field_backward_kernel": "ngp_field_backward", "train_ray_loss_kernel": "ngp_train_ray_loss",
                "march_slab_kernel": "ngp_march_rays_train", "march_packed_kernel": "ngp_march_rays_train_packed", "adam_step_fused_kernel": "ngp_adam_step_fused"}
    agg = collections.defaultdict(list)
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        for k, entry in entry_of.items():
            if k in name:
                rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
                wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
                ur, uw = rows[1][idx["dram__bytes_read.sum"]], rows[1][idx["dram__bytes_write.sum"]]
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                agg[entry].append((rd * mult.get(ur, 1) + wr * mult.get(uw, 1), float(r[idx["gpu__time_duration.sum"]].replace(",", ""))))
    res = {}
    for entry, vals in agg.items():
        # launches of very different sizes (e.g. the 2 M-point occupancy refresh) would blur the mean: keep the modal half
        vals.sort()
        mid = vals[len(vals) // 4: max(len(vals) // 4 + 1, 3 * len(vals) // 4)]
        res[entry] = {"dram_bytes_per_launch": sum(v[0] for v in mid) / len(mid), "launches_in_capture": len(vals),
                      "samples_per_launch": samples_per_launch, "capture": src.split("/")[-1],
                      "how": "ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum, mean of the "
                             "middle half of the captured launches"}
    with open(dst, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], float(sys.argv[4]) if len(sys.argv) > 4 else None)
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        raw(sys.argv[2], sys.argv[3])
