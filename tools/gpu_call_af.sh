#!/bin/bash
# Round-2 GPU call AF (1 GPU): tile shapes of k_mg_dir_apply / k_mg_update, rows per block of k_predict_first / k_mg_init (A/B hooks).
out=gpurun_out/r2af; mkdir -p $out
T="tests/test_gpu_baseline_configs.py::test_mgcg_complete_state_after_elided_recorrection_rounds tests/test_gpu_parity.py::test_mode_c_mgcg_matches_oracle_to_tolerance tests/test_gpu_parity.py::test_parabolic_first_order_no_cylinder tests/test_gpu_parity.py::test_cavity_extension_bit_exact"
n=0
for v in "CFD_DIR_TILE=2x128 CFD_UPD_TILE=2x128 CFD_INIT_ROWS=4 CFD_PRED_ROWS=8" "CFD_DIR_TILE=2x256 CFD_UPD_TILE=2x256 CFD_INIT_ROWS=16 CFD_PRED_ROWS=32"; do
  n=$((n+1)); env $v timeout 200 python -m pytest $T -m gpu -q -x > "$out/pytest_$n.txt" 2>&1; echo "$v: pytest rc=$? $(tail -1 "$out/pytest_$n.txt")"
done
export CFD_BENCH_NO_EXTRAS=1
one() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_$name.json 2> $out/bench_$name.err; echo "bench $name rc=$?"
  python - "$out/bench_$name.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ('ms_per_step','cg_iterations_per_step','step_frac_of_peak')}, 'e2e', d['e2e']['ms_per_step'])
PY
}
one default CFD_X=0
one dir2x256 CFD_DIR_TILE=2x256
one dir2x128 CFD_DIR_TILE=2x128
one upd2x256 CFD_UPD_TILE=2x256
one upd2x128 CFD_UPD_TILE=2x128
one pred8 CFD_PRED_ROWS=8
one pred32 CFD_PRED_ROWS=32
one init4 CFD_INIT_ROWS=4
one init16 CFD_INIT_ROWS=16
one default2 CFD_X=0
