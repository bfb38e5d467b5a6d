#!/bin/bash
# Round-2 GPU call G (1 GPU): GPU tests, bench, launch list; A/B of the two fused pre-smoothing kernels.
out=gpurun_out/r2g; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -k "not config1 and not config3" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -8 $out/pytest.txt
export CFD_BENCH_NO_EXTRAS=1
timeout 300 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
CFD_FUSED0_SIMPLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_simple0.json 2> $out/bench_simple0.err; echo "bench simple0 rc=$?"
for f in $out/bench.json $out/bench_simple0.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ('ms_per_step','cg_iterations_per_step','ms_per_cg_iteration','step_frac_of_peak','step_frac_of_peak_counting_elided_passes')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_blocking_get_snapshot'])
PY
done
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_short.json 2>&1 &&
CFD_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
  --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
echo "ncu rc=$?"
