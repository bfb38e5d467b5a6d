#!/bin/bash
# Round-2 GPU call C (1 GPU): GPU tests (without the three long BASELINE-config tests), short bench, launch list.
out=gpurun_out/r2c; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=8 -k "not config1 and not config3" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -15 $out/pytest.txt
CFD_BENCH_NO_EXTRAS=1 timeout 300 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
head -c 5000 $out/bench.json; tail -5 $out/bench.err
CFD_BENCH_NO_EXTRAS=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_short.json 2>&1 &&
CFD_BENCH_NO_EXTRAS=1 CFD_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
  --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
echo "ncu rc=$?"
