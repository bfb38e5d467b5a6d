#!/bin/bash
# Round-2 GPU call E (2 GPUs): GPU tests incl. the strips test, the strips check over the peer-memory layer, 2-GPU bench lines.
out=gpurun_out/r2e; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=8 -k "not config1 and not config3" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -12 $out/pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
mkdir -p $out/logs_peer
CFD_PEER_STRIPS=1 CFD_STRIP_LOG_DIR=$out/logs_peer timeout 600 $TR --master-port 29621 tests/mgpu_strip_check.py > $out/strips_peer.txt 2>&1; echo "strips peer rc=$?"
tail -5 $out/strips_peer.txt; tail -3 $out/logs_peer/strip_check_rank0.log
export CFD_BENCH_NO_EXTRAS=1
timeout 600 $TR --master-port 29622 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2_nccl.json 2> $out/bench_n2_nccl.err; echo "bench n2 nccl rc=$?"
CFD_PEER_STRIPS=1 timeout 600 $TR --master-port 29623 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2_peer.json 2> $out/bench_n2_peer.err; echo "bench n2 peer rc=$?"
timeout 600 $TR --master-port 29624 bench.py --gpus 2 --steps 3 --warmup 3 --workload channel8192x2048_modeR > $out/bench_ch_n2_nccl.json 2> $out/bench_ch_n2_nccl.err; echo "bench ch n2 nccl rc=$?"
CFD_PEER_STRIPS=1 timeout 600 $TR --master-port 29625 bench.py --gpus 2 --steps 3 --warmup 3 --workload channel8192x2048_modeR > $out/bench_ch_n2_peer.json 2> $out/bench_ch_n2_peer.err; echo "bench ch n2 peer rc=$?"
for f in $out/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('ms_per_step','value','cg_iterations_per_step','sweeps_per_step','ms_per_cg_iteration')}, d.get('e2e',{}).get('ms_per_step'), json.dumps(d.get('parity'))[:900])
except Exception as e:
    print('no line', e)
PY
done
tail -3 $out/*.err | tail -30
