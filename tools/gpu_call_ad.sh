#!/bin/bash
# Round-2 GPU call AD (1 GPU): corrector fused with the re-correction round's divergence (k_corrector_div), dot-product
# finishes of k_mg_init / k_mg_dir_apply / k_mg_update as per-block partials + k_mg_reduce — parity tests, A/B benches, launch list.
out=gpurun_out/r2ad; mkdir -p $out
K="mgcg or mode_c or legs or relative or headline or deterministic"
timeout 600 python -m pytest tests -m gpu -q --maxfail=8 --durations=3 -k "$K" > $out/pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc" >> $out/pytest.txt
tail -4 $out/pytest.txt
if [ $rc -ne 0 ]; then
  grep -E "^(FAILED|ERROR)" $out/pytest.txt | head -12
  CFD_FUSED_CORRECTOR=0 timeout 300 python -m pytest tests -m gpu -q --lf --maxfail=8 > $out/pytest_nofuse.txt 2>&1; echo "no fused corrector: rc=$?"; tail -2 $out/pytest_nofuse.txt
  CFD_MG_FINISH_LAUNCH=0 timeout 300 python -m pytest tests -m gpu -q --lf --maxfail=8 > $out/pytest_nofinish.txt 2>&1; echo "ticket finish: rc=$?"; tail -2 $out/pytest_nofinish.txt
fi
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
one nofuse CFD_FUSED_CORRECTOR=0
one ticket CFD_MG_FINISH_LAUNCH=0
one rows8 CFD_CORR_ROWS=8
one default2 CFD_X=0
CFD_BENCH_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
  --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
python tools/launch_list.py $out/launches.csv "r2 call AD, fused corrector + divergence, finish launches" > $out/launch_list.txt 2>&1; head -16 $out/launch_list.txt
