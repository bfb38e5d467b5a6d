#!/bin/bash
# Round-2 GPU call J (1 GPU): generalized V(nu,nu) legs — parity tests, bench at nu = 2, 3, 4, ncu --set full of the level-0 legs.
out=gpurun_out/r2j; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -k "mgcg or mode_c or legs or relative" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -12 $out/pytest.txt
export CFD_BENCH_NO_EXTRAS=1
for nu in 2 3 4; do
CFD_BENCH_CONSTS="mg_smoothing=$nu" timeout 300 python bench.py --steps 20 --warmup 5 > $out/bench_nu$nu.json 2> $out/bench_nu$nu.err; echo "bench nu=$nu rc=$?"
python - "$out/bench_nu$nu.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ('ms_per_step','cg_iterations_per_step','ms_per_cg_iteration','step_frac_of_peak')}, 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['avg_launch_us'], d['cg_iterations_list'], d['stop']['rel_residual'], d['cpu_baseline'] and d['cpu_baseline']['sample'])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_mg0_(up|down)' -c 2 -o $out/legs0 python tools/profile_mg.py cavity4096_modeC 112 > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/legs0.ncu-rep --page raw --csv > $out/legs0_raw.csv 2>/dev/null; python tools/ncu_summary.py $out/legs0_raw.csv > $out/legs0_summary.txt 2>&1; cat $out/legs0_summary.txt
