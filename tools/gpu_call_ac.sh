#!/bin/bash
# Round-2 GPU call AC (1 GPU): divergence with per-block partials + reduce launch — parity tests, bench, launch list.
out=gpurun_out/r2ac; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=3 -k "mgcg or mode_c or legs or relative" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -4 $out/pytest.txt
export CFD_BENCH_NO_EXTRAS=1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - "$out/bench.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ('ms_per_step','cg_iterations_per_step','ms_per_cg_iteration','step_frac_of_peak','step_frac_of_peak_fused_traffic')}, 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['avg_launch_us'], d['roofline']['frac'])
PY
CFD_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
  --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
python tools/launch_list.py $out/launches.csv "r2 call AC, V(3,3), register-tiled legs, batched first-stage loads" > $out/launch_list.txt 2>&1; head -24 $out/launch_list.txt
