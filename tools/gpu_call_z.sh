#!/bin/bash
# Round-2 GPU call Z (1 GPU): the complete GPU test suite and the default bench line (with extras and the CPU leg).
out=gpurun_out/r2z; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --durations=8 > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -15 $out/pytest.txt
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - "$out/bench.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('ms_per_step','value','cg_iterations_per_step','ms_per_cg_iteration','step_frac_of_peak','gpu_launches')}, 'e2e', d['e2e'])
print('roofline', d['roofline']); print('cpu', d['cpu_baseline'])
for k,v in (d.get('extra') or {}).items(): print(k, {a:v.get(a) for a in ('ms_per_step','cell_updates_per_s','cg_iterations_per_step','sweeps_per_step','ms_per_cg_iteration','sweep_us','step_frac_of_peak','error')})
PY
tail -3 $out/bench.err
