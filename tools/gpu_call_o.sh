#!/bin/bash
# Round-2 GPU call O (1 GPU): ncu --set full of the register-tiled legs (level 0 and level 1) at V(3,3).
out=gpurun_out/r2o; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_mg0_(up|down)3|k_mgc_(up|down)3' -c 4 -o $out/legs3 python tools/profile_mg.py cavity4096_modeC 112 > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/legs3.ncu-rep --page raw --csv > $out/legs3_raw.csv 2>/dev/null; python tools/ncu_summary.py $out/legs3_raw.csv > $out/legs3_summary.txt 2>&1; cat $out/legs3_summary.txt | grep -v "^  launch__occ\|per_second\|hit_rate\|lg_throttle\|dispatch_stall\|branch_resolving"
