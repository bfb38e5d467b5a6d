#!/bin/bash
# Round-2 GPU call AG (1 GPU): validation of the session's build — default bench line (extras + CPU leg), smoke(), the GPU test suite.
out=gpurun_out/r2ag; mkdir -p $out
t0=$(date +%s)
timeout 240 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
python - "$out/bench.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('ms_per_step','value','cg_iterations_per_step','ms_per_cg_iteration','step_frac_of_peak','step_frac_of_peak_fused_traffic','gpu_launches')}, 'e2e', d['e2e'])
print('roofline', d['roofline']); print('cpu', d['cpu_baseline'])
for k,v in (d.get('extra') or {}).items(): print(k, {a:v.get(a) for a in ('ms_per_step','cell_updates_per_s','cg_iterations_per_step','sweeps_per_step','ms_per_cg_iteration','sweep_us','step_frac_of_peak','error')})
PY
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.txt
timeout 330 python -m pytest tests -m gpu -q --maxfail=20 --durations=5 --deselect tests/test_gpu_baseline_configs.py::test_config3_channel_8192x2048_saturated_step_bit_exact > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -9 $out/pytest.txt
echo "elapsed $(( $(date +%s) - t0 )) s"
timeout 130 python -m pytest "tests/test_gpu_baseline_configs.py::test_config3_channel_8192x2048_saturated_step_bit_exact" -m gpu -q > $out/pytest_channel.txt 2>&1; echo "channel pytest rc=$?"; tail -2 $out/pytest_channel.txt
