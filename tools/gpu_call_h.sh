#!/bin/bash
# Round-2 GPU call H (1 GPU): GPU tests (persistent small-grid solve, table sweeps), Mode R workloads, default bench.
out=gpurun_out/r2h; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -k "not config1_cavity1024_mode_c and not config3" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -8 $out/pytest.txt
export CFD_BENCH_NO_EXTRAS=1
for wl in default800_modeR cavity1024_modeR; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl > $out/bench_$wl.json 2> $out/bench_$wl.err; echo "bench $wl rc=$?"
  CFD_BENCH_FLAGS=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl > $out/bench_${wl}_nopersist.json 2> $out/bench_${wl}_nopersist.err; echo "bench $wl nopersist rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
for f in $out/bench*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ('ms_per_step','sweeps_per_step','cg_iterations_per_step','step_frac_of_peak')}, 'sweep_us', d['roofline']['avg_launch_us'], 'e2e', d['e2e']['ms_per_step'])
PY
done
