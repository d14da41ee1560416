"""DRAM traffic per launch of every entry point of the training step, from an
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --csv` pass over
profiles/train_step_eager.py (the kernels of a step, serialised).  bench.py reads the JSON this writes for
`roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the entry point's launches).
Usage: python profiles/make_traffic.py gpurun_out/metrics.csv <steps in the capture> profiles/r01_dram_traffic.json"""
import collections
import csv
import json
import re
import sys

# kernel-name pattern -> C-ABI entry point whose launches it implements in the TRAINING step
# (linear_tc*: the statistics epilogue <1>/<true> is the forward layer, <0>/<false> the data gradient)
ENTRY = [
    (r"bwd_fused_kernel", "pn2_mlp_bwd_layer"),
    (r"bg_(build|query)_kernel", "pn2_query_ball_point_grid"),
    (r"wgrad_tc_kernel", "pn2_linear_bwd_weight_accum"),
    (r"linear_tc2?_kernel<(1|true)>", "pn2_linear_fwd_prepacked"),
    (r"linear_tc2?_kernel<(0|false)>", "pn2_linear_bwd_data_prepacked"),
    (r"pool_bwd_dz_vec8_kernel", "pn2_pool_bn_relu_bwd_dz"),
    (r"bn_bwd_dz", "pn2_bn_relu_bwd_dz"),
    (r"bn_bwd_reduce_vec8_kernel<2>", "pn2_pool_bn_relu_bwd_reduce_finalize"),
    (r"bn_bwd_reduce", "pn2_bn_relu_bwd_reduce_finalize"),
    (r"bn_relu_max", "pn2_bn_relu_max_keep"),
    (r"bn_relu_kernel", "pn2_bn_relu"),
    (r"fps_kernel", "pn2_farthest_point_sample"),
    (r"ball_query_kernel", "pn2_query_ball_point"),
    (r"three_nn_kernel", "pn2_three_nn"),
    (r"interp_concat_kernel", "pn2_interp_concat"),
    (r"interp_bwd_kernel", "pn2_interp_bwd"),
    (r"group_points_bwd", "pn2_group_points_bwd"),
    (r"group_points", "pn2_group_points"),
    (r"head_tail_fwd_kernel", "pn2_head_tail_loss_fwd"),       # the training step calls the tail fused with the loss
    (r"head_tail_bwd_kernel", "pn2_head_tail_loss_bwd"),
    (r"adam_flat_kernel", "pn2_adam_step"),
    (r"rows_to_f32_kernel", "pn2_rows_to_f32"),
    (r"pack_weights", "pn2_pack_weights"),
]
UNIT = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, steps, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        except Exception:
            continue
        per.setdefault((row["ID"], row["Kernel Name"]), {})[row["Metric Name"]] = v
    agg = {}
    for (_, name), m in per.items():
        entry = next((e for pat, e in ENTRY if re.search(pat, name)), None)
        if entry is None:
            continue
        a = agg.setdefault(entry, {"launches": 0, "bytes": 0.0, "seconds": 0.0})
        a["launches"] += 1
        a["bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a["seconds"] += m.get("gpu__time_duration.sum", 0.0)
    res = {"_source": "%s (%d eager steps of profiles/train_step_eager.py under ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
                      "gpu__time_duration.sum --clock-control none)" % (path, steps)}
    for e, a in sorted(agg.items(), key=lambda kv: -kv[1]["seconds"]):
        res[e] = {"launches_per_step": a["launches"] / steps, "dram_bytes_per_step": a["bytes"] / steps,
                  "dram_bytes_per_launch": a["bytes"] / a["launches"],
                  "ncu_us_per_launch": a["seconds"] * 1e6 / a["launches"]}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", out, "with", len(res) - 1, "entry points")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3])
