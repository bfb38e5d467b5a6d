#!/bin/bash
# Round-2 GPU call AH (1 GPU): the default bench line with the multi-pass e2e region.
out=gpurun_out/r2ah; mkdir -p $out
timeout 80 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - "$out/bench.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('ms_per_step','value','cg_iterations_per_step','step_frac_of_peak','step_frac_of_peak_fused_traffic','gpu_launches')})
e=d['e2e']; print('e2e', e['ms_per_step'], e['ms_per_step_passes'], e['ms_per_step_median_pass'], e['ms_per_step_blocking_get_snapshot'])
PY
tail -3 $out/bench.err
