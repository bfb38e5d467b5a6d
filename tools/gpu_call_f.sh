#!/bin/bash
# Round-2 GPU call F (2 GPUs): GPU tests (rolling predictor etc.), strips check over NCCL and over the peer-memory layer, 2-GPU bench.
out=gpurun_out/r2f; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -k "not config1 and not config3" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -8 $out/pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
mkdir -p $out/logs_peer
CFD_PEER_STRIPS=1 CFD_STRIP_CHECK_SMALL=1 CFD_STRIP_LOG_DIR=$out/logs_peer timeout 600 $TR --master-port 29631 tests/mgpu_strip_check.py > $out/strips_peer.txt 2>&1; echo "strips peer rc=$?"
tail -3 $out/strips_peer.txt; tail -4 $out/logs_peer/strip_check_rank0.log
export CFD_BENCH_NO_EXTRAS=1
CFD_BENCH_NO_EXTRAS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench n1 rc=$?"
timeout 600 $TR --master-port 29632 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2_nccl.json 2> $out/bench_n2_nccl.err; echo "bench n2 nccl rc=$?"
CFD_PEER_STRIPS=1 timeout 600 $TR --master-port 29633 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2_peer.json 2> $out/bench_n2_peer.err; echo "bench n2 peer rc=$?"
CFD_PEER_STRIPS=1 timeout 600 $TR --master-port 29635 bench.py --gpus 2 --steps 3 --warmup 3 --workload channel8192x2048_modeR > $out/bench_ch_n2_peer.json 2> $out/bench_ch_n2_peer.err; echo "bench ch n2 peer rc=$?"
for f in $out/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('ms_per_step','value','cg_iterations_per_step','sweeps_per_step','ms_per_cg_iteration','step_frac_of_peak')}, d.get('e2e',{}).get('ms_per_step'), json.dumps(d.get('parity'))[:400])
except Exception as e:
    print('no line', e)
PY
done
for f in $out/*.err; do tail -n 3 $f; done
