#!/bin/bash
# Round-2 GPU call B (1 GPU): full GPU test suite, default bench line (with extras), launch list of the same bench.
out=gpurun_out/r2b; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --durations=12 > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -25 $out/pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
head -c 9000 $out/bench.json; tail -5 $out/bench.err
CFD_BENCH_NO_EXTRAS=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_short.json 2>&1 &&
CFD_BENCH_NO_EXTRAS=1 CFD_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
  --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
echo "ncu rc=$?"
