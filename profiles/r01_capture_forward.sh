#!/bin/bash
# Round-1 profile capture of the INFERENCE forward (eager, one 32 x 4096 batch): ncu metric pass over every launch and
# --set full of the fused set-abstraction kernel (K3d) and the feature-propagation kernels; then the other bench configs.
set -x
CMD="python profiles/forward_only.py"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
$CMD > gpurun_out/r01_fwd_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r01_fwd_plain.log; exit 1; }
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r01_metrics_forward.csv $CMD > gpurun_out/r01_ncu_fwd_metrics.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:sa_fused_eval_kernel|interp_concat_kernel" -s 8 -c 8 -o gpurun_out/r01_full_forward -f $CMD > gpurun_out/r01_ncu_fwd_full.log 2>&1
ncu -i gpurun_out/r01_full_forward.ncu-rep --page raw --csv > gpurun_out/r01_full_forward.raw.csv 2>/dev/null
python bench.py --workload fps_ball --warmup 3 > gpurun_out/r01_bench_fps_ball_n1.json 2> gpurun_out/r01_bench_fps_ball_n1.err
python bench.py --channels 6 --no-cpu-baseline > gpurun_out/r01_bench_train_6ch_n1.json 2> gpurun_out/r01_bench_train_6ch_n1.err
python bench.py --workload facade --from-scene --warmup 3 > gpurun_out/r01_bench_facade_from_scene_n1.json 2> gpurun_out/r01_bench_facade_from_scene_n1.err
ls -la gpurun_out/r01_*forward* gpurun_out/r01_bench_*
