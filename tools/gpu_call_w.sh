#!/bin/bash
# Round-2 GPU call W (2 GPUs): strips with the register-tiled legs at V(3,3) after the even-first-row fix — strip check over
# NCCL and over peer memory, weak-scaling bench at N = 2 (both transports).
out=gpurun_out/r2w; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export CFD_STRIP_CHECK_SMALL=1 CFD_STRIP_LOG_DIR=$out
CFD_STRIP_CHECK_NU=3 timeout 600 $TR --master-port 29711 tests/mgpu_strip_check.py > $out/strip_nccl_nu3.txt 2>&1; echo "strip check nccl nu3 rc=$?"; tail -2 $out/strip_nccl_nu3.txt
CFD_PEER_STRIPS=1 CFD_STRIP_CHECK_NU=3 timeout 600 $TR --master-port 29712 tests/mgpu_strip_check.py > $out/strip_peer_nu3.txt 2>&1; echo "strip check peer nu3 rc=$?"; tail -2 $out/strip_peer_nu3.txt
export CFD_BENCH_NO_EXTRAS=1
CFD_PEER_STRIPS=1 timeout 600 $TR --master-port 29715 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2_peer.json 2> $out/bench_n2_peer.err; echo "bench n2 peer rc=$?"
timeout 600 $TR --master-port 29716 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2_nccl.json 2> $out/bench_n2_nccl.err; echo "bench n2 nccl rc=$?"
for f in $out/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('ms_per_step','value','cg_iterations_per_step','ms_per_cg_iteration','gpu_launches')}, 'e2e', d.get('e2e',{}).get('ms_per_step'))
    print(json.dumps(d.get('parity'))[:900])
except Exception as e:
    print('no line', e)
PY
done
for f in $out/*.err; do tail -n 3 $f | cut -c1-300; done
