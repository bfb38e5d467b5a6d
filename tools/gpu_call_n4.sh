#!/bin/bash
# Round-2 GPU call N4 (4 GPUs): strip check (small cases, V(3,3)) over NCCL and the default bench line at N = 4 with its extras.
out=gpurun_out/r2n4; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
CFD_STRIP_CHECK_SMALL=1 CFD_STRIP_LOG_DIR=$out timeout 600 $TR --master-port 29711 tests/mgpu_strip_check.py > $out/strip_nccl.txt 2>&1; echo "strip check nccl rc=$?"; tail -2 $out/strip_nccl.txt
timeout 900 $TR --master-port 29652 bench.py --gpus 4 --steps 10 --warmup 3 > $out/bench_n4.json 2> $out/bench_n4.err; echo "bench n4 rc=$?"
python - "$out/bench_n4.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('ms_per_step','value','cg_iterations_per_step','ms_per_cg_iteration','step_frac_of_peak','gpu_launches')}, 'e2e', d.get('e2e',{}).get('ms_per_step'))
    print(json.dumps(d.get('parity'))[:900])
    for k,v in (d.get('extra') or {}).items(): print(k, {a:v.get(a) for a in ('ms_per_step','cell_updates_per_s','cg_iterations_per_step','sweeps_per_step','error')})
except Exception as e:
    print('no line', e)
PY
tail -n 4 $out/bench_n4.err | cut -c1-300
