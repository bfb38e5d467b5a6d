#!/bin/bash
# Round-2 GPU call AI (1 GPU): ncu --set full of k_corrector_div (after its own command exited 0 in call AH), then the launch list.
out=gpurun_out/r2ai; mkdir -p $out
export CFD_BENCH_NO_EXTRAS=1
CFD_BENCH_PROFILE=1 timeout 60 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_corrector_div' -c 1 \
  -o $out/corrector_div python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/corrector_div.ncu-rep --page raw --csv > $out/corrector_div_raw.csv 2>/dev/null; python tools/ncu_summary.py $out/corrector_div_raw.csv > $out/corrector_div_summary.txt 2>&1; head -34 $out/corrector_div_summary.txt
CFD_BENCH_PROFILE=1 timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
  --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1; echo "ncu list rc=$?"
python tools/launch_list.py $out/launches.csv "r2 call AI, end of round 2: k_corrector_div 2 x 128, 4-row divergence, 2-row direction kernel, 8-row predictor blocks" > $out/launch_list.txt 2>&1; head -22 $out/launch_list.txt
