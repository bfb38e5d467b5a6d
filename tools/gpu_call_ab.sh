#!/bin/bash
# Round-2 GPU call AB (1 GPU): k_predict_first with the masks loaded one row ahead — GPU tests (all but the three long ones), launch list.
out=gpurun_out/r2ab; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=3 -k "not config1 and not config3" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -4 $out/pytest.txt
export CFD_BENCH_NO_EXTRAS=1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - "$out/bench.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('ms_per_step','cg_iterations_per_step','step_frac_of_peak','step_frac_of_peak_fused_traffic')}, 'e2e', d['e2e']['ms_per_step'])
PY
CFD_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
  --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
python tools/launch_list.py $out/launches.csv "r2 call AB" > $out/launch_list.txt 2>&1; head -12 $out/launch_list.txt
