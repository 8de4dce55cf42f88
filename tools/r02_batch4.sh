#!/bin/bash
# Round 2, batch 4: config 3 at full size (basis path), config 4 full IPM on the device arm, ncu evidence.
out=gpurun_out
mkdir -p $out
( time timeout 1500 python tools/fullsize.py --basis C3 --scale 1.0 --volume-tol 1e30 --kkt-maxiter 100 --out $out/r02d_c3_full_basis.json > $out/r02d_c3.log 2>&1 ) 2>&1 | grep real; echo "c3 rc=$?"; tail -c 3000 $out/r02d_c3.log
( time timeout 900 python tools/solve_lp.py transport:2000:5000 --impl gpu --crossover 1 --out $out/r02d_c4_gpu.json > $out/r02d_c4_gpu.log 2>&1 ) 2>&1 | grep real; tail -2 $out/r02d_c4_gpu.log | cut -c1-1200
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 --no-ipm --no-parity"
$B > $out/r02d_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches_bench.csv $B > $out/r02d_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
P="python tools/profile_apply.py --reps 2 --pcr 6"
$P > $out/r02d_plain_profile.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'band_sweep|pcr_fused|band_combine' -c 8 -o $out/r02_prof_band_fused $P > $out/r02d_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/r02_prof_band_fused.ncu-rep --page raw --csv > $out/r02_ncu_full_band_fused.csv 2>/dev/null; ls -la $out | tail -12
