#!/bin/bash
# Round-1 profile capture (run under gpurun on one B200): plain run first, then
#  (1) ncu metrics pass over 2 eager training steps: duration + DRAM bytes of every launch  -> gpurun_out/r01_metrics_train.csv
#  (2) ncu --set full of representative launches of the top kernel families                 -> gpurun_out/r01_full_*.ncu-rep (+ raw csv)
#  (3) torch.profiler timeline of one graph replay (pipelined train step and forward)       -> gpurun_out/r01_trace_*.csv
set -x
CMD="python profiles/train_step_eager.py 2"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
$CMD > gpurun_out/r01_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r01_plain.log; exit 1; }
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r01_metrics_train.csv $CMD > gpurun_out/r01_ncu_metrics.log 2>&1
full() {   # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o gpurun_out/r01_full_$1 -f $CMD > gpurun_out/r01_ncu_$1.log 2>&1
  ncu -i gpurun_out/r01_full_$1.ncu-rep --page raw --csv > gpurun_out/r01_full_$1.raw.csv 2>/dev/null
}
full linear_fwd_sa1 'linear_tc_kernel' 0 3
full wgrad_sa1 'wgrad_tc_kernel' 20 3
full bn_bwd_sa1 'bn_bwd_(dz|reduce)_vec8' 32 4
full linear_tc2_fp1 'linear_tc2_kernel' 14 4
full geometry 'fps_kernel|ball_query_kernel|three_nn_kernel' 0 3
full pool_tail 'pool_bwd_dz|bn_relu_max' 0 2
full head_adam 'head_tail_fwd_kernel|head_tail_bwd_kernel|adam_flat_kernel' 0 3
python profiles/trace_step.py train gpurun_out/r01_trace_train_pipelined.csv
python profiles/trace_step.py forward gpurun_out/r01_trace_forward_pipelined.csv
ls -la gpurun_out/r01_*
