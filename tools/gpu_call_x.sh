#!/bin/bash
# Round-2 GPU call X (8 GPUs): the default bench line at N = 8 over NCCL with the batched exchanges (weak 4096^2 per GPU; extras:
# 16384^2 and 8192x2048 strong).
out=gpurun_out/r2x; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29652 bench.py --gpus 8 --steps 10 --warmup 3 > $out/bench_n8_nccl.json 2> $out/bench_n8_nccl.err; echo "bench n8 nccl rc=$?"
for f in $out/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('ms_per_step','value','cg_iterations_per_step','ms_per_cg_iteration','step_frac_of_peak','gpu_launches')}, 'e2e', d.get('e2e',{}).get('ms_per_step'))
    print(json.dumps(d.get('parity'))[:900])
    for k,v in (d.get('extra') or {}).items(): print(k, {a:v.get(a) for a in ('ms_per_step','cell_updates_per_s','cg_iterations_per_step','sweeps_per_step','ms_per_cg_iteration','sweep_us','step_frac_of_peak','error')})
except Exception as e:
    print('no line', e)
PY
done
tail -n 4 $out/bench_n8_nccl.err | cut -c1-300
