#!/bin/bash
# Round-2 GPU call L (1 GPU): ncu --set full of the column-strip level-0 legs and the level-1 legs at V(3,3).
out=gpurun_out/r2l; mkdir -p $out
export CFD_BENCH_CONSTS="mg_smoothing=3"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_mg0_(up|down)2|k_mgc_(up|down)' -c 4 -o $out/legs python tools/profile_mg.py cavity4096_modeC 112 > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/legs.ncu-rep --page raw --csv > $out/legs_raw.csv 2>/dev/null; python tools/ncu_summary.py $out/legs_raw.csv > $out/legs_summary.txt 2>&1; cat $out/legs_summary.txt | grep -v "^  launch__occ\|per_second"
