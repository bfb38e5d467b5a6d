#!/bin/bash
# Round-2 GPU call A (1 GPU): full GPU test suite, default bench line, launch list of the same bench.
out=gpurun_out/r2a; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --durations=20 > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -40 $out/pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cat $out/bench.json | head -c 6000
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_short.json 2>&1 &&
CFD_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
  --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
echo "ncu rc=$?"
