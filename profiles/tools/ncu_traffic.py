#!/usr/bin/env python
"""DRAM traffic per launch of the hot kernels, from an `ncu --set full` capture -> profiles/ncu_traffic.json (read by
bench.py for `roofline.traffic`).

    python profiles/tools/ncu_traffic.py gpurun_out/prof_r2f2.ncu-rep <samples per launch of the captured run> [out.json]

Per kernel: dram__bytes_read.sum + dram__bytes_write.sum of its FIRST captured launch (the capture is taken on eager
single-chain steps of the default bench workload, so one launch = one step's samples; consecutive steps render different
view batches with different sample counts, so launches are not averaged: all first launches belong to the same step)."""
import collections
import csv
import json
import os
import subprocess
import sys

ENTRY_OF = {
    "encode_backward_warpagg_kernel": "ngp_grid_scatter_samples", "field_forward_kernel": "ngp_field_forward",
    "field_backward_kernel": "ngp_field_backward", "march_packed_kernel": "ngp_march_rays_train_packed",
    "march_slab_kernel": "ngp_march_rays_train", "train_ray_loss_kernel": "ngp_train_ray_loss",
    "adam_step_kernel": "ngp_adam_step", "adam_step_fused_kernel": "ngp_adam_step_fused",
    "check_finite_fold_kernel": "ngp_check_finite_fold", "quad_table_kernel": "ngp_grid_quad_table",
}


def main():
    rep, samples = sys.argv[1], float(sys.argv[2])
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ncu_traffic.json")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                          "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in txt.splitlines() if l.startswith('"')]))
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    agg = collections.defaultdict(list)
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        entry = next((e for k, e in ENTRY_OF.items() if k in name), None)
        if entry is None:
            continue
        b = sum(float(r[col[m]].replace(",", "")) * scale[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        agg[entry].append((b, float(r[col["gpu__time_duration.sum"]].replace(",", ""))))
    res = {}
    for entry, v in sorted(agg.items()):
        res[entry] = {"capture": os.path.basename(rep), "dram_bytes_per_launch": round(v[0][0]),
                      "duration_us_under_ncu": round(v[0][1], 2), "launches_in_capture": len(v),
                      "samples_per_launch": samples,
                      "how": "ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum, first captured "
                             "launch of the kernel (eager single-chain step of the default bench workload)"}
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
