#!/bin/bash
# Round-2 GPU call T (1 GPU): largest shared-memory carve-out for the legs; 3 (80 registers) vs 4 (64 registers) blocks per SM.
out=gpurun_out/r2t; mkdir -p $out
export CFD_BENCH_NO_EXTRAS=1
for lib in shipped occ4; do
if [ $lib = occ4 ]; then export CFD_B200_LIB=$PWD/cfd_demo_b200/libcfd_b200_occ4.so; else unset CFD_B200_LIB; fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_$lib.json 2> $out/bench_$lib.err; echo "bench $lib rc=$?"
python - "$out/bench_$lib.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ('ms_per_step','cg_iterations_per_step','ms_per_cg_iteration','step_frac_of_peak','step_frac_of_peak_fused_traffic')}, 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['avg_launch_us'], d['roofline']['frac'])
PY
CFD_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
  --log-file $out/launches_$lib.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_$lib.log 2>&1
python tools/launch_list.py $out/launches_$lib.csv "r2 call T $lib" > $out/launch_list_$lib.txt 2>&1; head -5 $out/launch_list_$lib.txt
done
