#!/bin/bash
# Round-2 GPU call Y (1 GPU): k_jacobi_persist2 (grid barrier off the critical path) — GPU tests, default 800x264 grid A/B.
out=gpurun_out/r2y; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=3 -k "not config1 and not config3" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -5 $out/pytest.txt
for cfg in "1 8 1024" "2 8 1024" "2 4 1024" "2 2 1024" "2 16 1024" "2 8 512"; do set -- $cfg
CFD_PERSIST_FORM=$1 CFD_PERSIST_ROWS=$2 CFD_PERSIST_THREADS=$3 timeout 200 python bench.py --workload default800_modeR --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_f$1_r$2_t$3.json 2> $out/bench_f$1_r$2_t$3.err; echo "form $1 rows $2 threads $3 rc=$?"
python - "$out/bench_f$1_r$2_t$3.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('ms_per_step','sweeps_per_step','solves_per_step','gpu_launches')}, 'sweep_us', d['roofline']['avg_launch_us'])
except Exception as e: print('no line', e)
PY
done
