#!/bin/bash
# Round-2 GPU call AA (1 GPU): timing experiment — does the one-address atomicMax of 12 000 one-warp blocks cost the Mode R sweep
# its 15 us against the kDot form?  (CFD_SWEEP_SHARD_TEST=1: 63 of 64 blocks send their max to scratch slots.)
out=gpurun_out/r2aa; mkdir -p $out
for t in 0 1; do
if [ $t = 1 ]; then export CFD_SWEEP_SHARD_TEST=1; else unset CFD_SWEEP_SHARD_TEST; fi
timeout 300 python bench.py --workload channel4096_modeR --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_shard$t.json 2> $out/bench_shard$t.err; echo "shard_test=$t rc=$?"
python - "$out/bench_shard$t.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('ms_per_step','sweeps_per_step','solves_per_step','step_frac_of_peak')}, 'sweep_us', d['roofline']['avg_launch_us'], d['roofline']['frac'])
except Exception as e: print('no line', e)
PY
done
