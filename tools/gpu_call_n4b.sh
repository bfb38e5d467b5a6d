#!/bin/bash
# Round-2 GPU call N4b (4 GPUs): the pytest multi-GPU test as a 4-GPU box would run it (strip check, all cases that fit), smoke().
out=gpurun_out/r2n4b; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.txt
CFD_STRIP_LOG_DIR=$out timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "strips_over_nccl" > $out/pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest.txt
