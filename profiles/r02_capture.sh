#!/bin/bash
# Round-2 profile capture (run under gpurun on one B200; every profiled command first runs plain and must exit 0):
#  (1) ncu launch list of the bench command itself (gpu__time_duration.sum of every launch)     -> gpurun_out/r02_bench_launches.csv
#  (2) ncu metrics pass over 2 eager training steps: duration + DRAM bytes of every launch       -> gpurun_out/r02_metrics_train.csv
#  (3) ncu --set full of the kernels this round touched (fused backward layer, forward layer after the uniform-issue
#      change, cell-grid ball query)                                                             -> gpurun_out/r02_full_*.ncu-rep
#  (4) CUPTI timeline of two back-to-back graph replays of the pipelined step                    -> gpurun_out/r02_trace_train_pipelined.csv
set -x
BENCH="python bench.py --headline-only --no-cpu-baseline --steps 4 --warmup 3"
timeout 300 $BENCH > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/r02_bench_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_bench_launches.csv $BENCH > gpurun_out/r02_ncu_bench.log 2>&1
CMD="python profiles/train_step_eager.py 2"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
timeout 120 $CMD > gpurun_out/r02_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_plain.log; exit 1; }
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_metrics_train.csv $CMD > gpurun_out/r02_ncu_metrics.log 2>&1
full() {   # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o gpurun_out/r02_full_$1 -f $CMD > gpurun_out/r02_ncu_$1.log 2>&1
  ncu -i gpurun_out/r02_full_$1.ncu-rep --page raw --csv > gpurun_out/r02_full_$1.raw.csv 2>/dev/null
}
full bwd_fused 'bwd_fused_kernel' 11 11
full linear_fwd_sa1 'linear_tc_kernel' 0 3
full ball_grid 'bg_(build|query)_kernel' 0 2
full wgrad 'wgrad_tc_kernel' 10 2
timeout 300 python profiles/trace_step.py train gpurun_out/r02_trace_train_pipelined.csv
ls -la gpurun_out/r02_*
