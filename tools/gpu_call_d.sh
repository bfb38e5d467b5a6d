#!/bin/bash
# Round-2 GPU call D (1 GPU): ncu --set full of the elementwise / V-cycle kernels that sit furthest below the HBM roofline.
out=gpurun_out/r2d; mkdir -p $out
export CFD_BENCH_NO_EXTRAS=1
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/bench_short.json 2>&1 &&
CFD_BENCH_PROFILE=1 timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k 'regex:k_predict_first|k_mg_dir_apply|k_mg_fused_sweep|k_corrector|k_mg_fine_restrict|k_mgc_sweep|k_mg_init|k_divergence|k_mg_update' -c 16 \
  -o $out/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $out/ncu.log; ls -la $out
