#!/bin/bash
# Round-2 GPU call S (1 GPU): uniform-tile fast path of the coarse legs — parity tests, bench (shipped library vs the
# 4-blocks-per-SM build of the level-0 legs), launch list, ncu --set full of the ascending level-0 leg (traffic figure).
out=gpurun_out/r2s; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=5 -k "mgcg or mode_c or legs or relative" > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt
tail -6 $out/pytest.txt
export CFD_BENCH_NO_EXTRAS=1
for lib in shipped occ4; do
if [ $lib = occ4 ]; then export CFD_B200_LIB=$PWD/cfd_demo_b200/libcfd_b200_occ4.so; else unset CFD_B200_LIB; fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_$lib.json 2> $out/bench_$lib.err; echo "bench $lib rc=$?"
python - "$out/bench_$lib.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ('ms_per_step','cg_iterations_per_step','ms_per_cg_iteration','step_frac_of_peak','step_frac_of_peak_fused_traffic')}, 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['avg_launch_us'], d['roofline']['frac'])
PY
done
unset CFD_B200_LIB
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_short.json 2>&1 &&
CFD_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
  --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
echo "ncu rc=$?"
python tools/launch_list.py $out/launches.csv "r2 call S, V(3,3), register-tiled legs, uniform-tile fast path on the coarse levels" > $out/launch_list.txt 2>&1; head -24 $out/launch_list.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_mg0_up3' -c 1 -o $out/up3 python tools/profile_mg.py cavity4096_modeC 112 > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/up3.ncu-rep --page raw --csv > $out/up3_raw.csv 2>/dev/null; python tools/ncu_summary.py $out/up3_raw.csv > $out/up3_summary.txt 2>&1; head -12 $out/up3_summary.txt
