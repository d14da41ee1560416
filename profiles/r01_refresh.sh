#!/bin/bash
# Round-1 refresh after the fused head / flat Adam / two-slot pipeline changes: the parts of r01_capture.sh and
# r01_capture_forward.sh whose results changed (the tcgen05 layer, weight-gradient, BatchNorm and index kernels are the
# ones captured before; their --set full reports stay valid).
set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
CMD="python profiles/train_step_eager.py 2"
$CMD > gpurun_out/r01_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r01_plain.log; exit 1; }
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r01_metrics_train.csv $CMD > gpurun_out/r01_ncu_metrics.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:head_tail_fwd_kernel|head_tail_bwd_kernel|adam_flat_kernel" -c 3 -o gpurun_out/r01_full_head_adam -f $CMD > gpurun_out/r01_ncu_head_adam.log 2>&1
ncu -i gpurun_out/r01_full_head_adam.ncu-rep --page raw --csv > gpurun_out/r01_full_head_adam.raw.csv 2>/dev/null
FWD="python profiles/forward_only.py"
$FWD > gpurun_out/r01_fwd_plain.log 2>&1 && ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r01_metrics_forward.csv $FWD > gpurun_out/r01_ncu_fwd_metrics.log 2>&1
python bench.py > gpurun_out/r01_bench_train_n1.json 2> gpurun_out/r01_bench_train_n1.err
python bench.py --workload fps_ball --warmup 3 > gpurun_out/r01_bench_fps_ball_n1.json 2> gpurun_out/r01_bench_fps_ball_n1.err
python bench.py --channels 6 --no-cpu-baseline > gpurun_out/r01_bench_train_6ch_n1.json 2> gpurun_out/r01_bench_train_6ch_n1.err
python bench.py --workload facade --warmup 3 > gpurun_out/r01_bench_facade_n1.json 2> gpurun_out/r01_bench_facade_n1.err
python bench.py --workload facade --from-scene --warmup 3 > gpurun_out/r01_bench_facade_from_scene_n1.json 2> gpurun_out/r01_bench_facade_from_scene_n1.err
tail -c 600 gpurun_out/r01_bench_train_n1.err
ls -la gpurun_out/r01_*
