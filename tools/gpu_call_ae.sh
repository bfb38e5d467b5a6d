#!/bin/bash
# Round-2 GPU call AE (1 GPU): tile shapes of k_corrector_div and k_divergence (A/B through the environment hooks).
out=gpurun_out/r2ae; mkdir -p $out
T="tests/test_gpu_baseline_configs.py::test_mgcg_complete_state_after_elided_recorrection_rounds tests/test_gpu_parity.py::test_mode_c_mgcg_matches_oracle_to_tolerance"
for v in "CFD_CORR_ROWS=2" "CFD_CORR_ROWS=1" "CFD_CORR_ROWS=2 CFD_CORR_THREADS=128" "CFD_DIV_ROWS=4"; do
  env $v timeout 200 python -m pytest $T -m gpu -q -x > "$out/pytest_${v// /_}.txt" 2>&1; echo "$v: pytest rc=$? $(tail -1 "$out/pytest_${v// /_}.txt")"
done
export CFD_BENCH_NO_EXTRAS=1
one() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_$name.json 2> $out/bench_$name.err; echo "bench $name rc=$?"
  python - "$out/bench_$name.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ('ms_per_step','cg_iterations_per_step','step_frac_of_peak','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'])
PY
}
one default CFD_X=0
one rows2 CFD_CORR_ROWS=2
one rows2t128 CFD_CORR_ROWS=2 CFD_CORR_THREADS=128
one rows1 CFD_CORR_ROWS=1
one rows4t128 CFD_CORR_THREADS=128
one div4 CFD_DIV_ROWS=4
one default2 CFD_X=0
