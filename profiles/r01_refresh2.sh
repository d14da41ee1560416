#!/bin/bash
# Last pass of round 1 (after bn_bwd_dz_rows_kernel): metric pass over 2 eager train steps, --set full of the new dz kernel,
# timeline of two back-to-back replays.
set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
CMD="python profiles/train_step_eager.py 2"
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r01_metrics_train.csv $CMD > gpurun_out/r01_ncu_metrics.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:bn_bwd_dz_rows_kernel" -s 8 -c 3 -o gpurun_out/r01_full_dz_rows -f python profiles/train_step_eager.py 1 > gpurun_out/r01_ncu_dz_rows.log 2>&1
ncu -i gpurun_out/r01_full_dz_rows.ncu-rep --page raw --csv > gpurun_out/r01_full_dz_rows.raw.csv 2>/dev/null
python profiles/trace_step.py train gpurun_out/r01_trace_train_pipelined.csv --replays=2
