run() { # name, env...
  n=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err || tail -5 gpurun_out/ab_$n.err
}
run p1f0 PN2_MAIN_PRIORITY=-1 PN2_TRAIN_FORK=0
run p1f1 PN2_MAIN_PRIORITY=-1 PN2_TRAIN_FORK=1
run p0f0 PN2_MAIN_PRIORITY=0 PN2_TRAIN_FORK=0
run p5f0 PN2_MAIN_PRIORITY=-5 PN2_TRAIN_FORK=0
python - <<'P'
import json
for f in ("p1f0","p1f1","p0f0","p5f0"):
    try:
        d=json.load(open("gpurun_out/ab_%s.json"%f)); print(f, round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), "fwd", round(d["forward"]["ms_per_batch"],4), round(d["forward"]["e2e_ms_per_batch"],4), "unpip", round(d["unpipelined"]["ms_per_step"],4))
    except Exception as e: print(f, "failed", e)
P
python profiles/trace_step.py forward gpurun_out/trace_forward_v24.csv --replays=2
python profiles/trace_step.py train gpurun_out/trace_train_v24.csv --replays=2
