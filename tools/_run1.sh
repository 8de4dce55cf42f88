B=./ipx_b200/_build/bsweep_bench
timeout 120 $B 3000 40000 10 t8,4,2,1024,2048,700,5,4 t16,3,2,1024,2048,700,5,2 > gpurun_out/bs_small.log 2>&1; echo "small rc=$?"; grep -E 'relerr|APPLY' gpurun_out/bs_small.log
timeout 300 $B 31,4,2,8192,8192,6250,14,0 31,4,2,8192,8192,6250,14,1 31,4,2,8192,8192,6250,14,2 31,4,2,8192,8192,6250,14,3 16,8,2,8192,8192,6250,14,1 > gpurun_out/bs_a.log 2>&1; echo "a rc=$?"
BSWEEP_DBG=1 timeout 300 $B t16,4,2,6144,6144,6250,19,3 t16,3,2,8192,8192,6250,14,2 t16,4,2,6144,6144,6250,19,2 t8,4,2,6144,6144,6250,19,6 t8,4,2,7168,7168,6250,16,4 > gpurun_out/bs_b.log 2>&1; echo "b rc=$?"
BSWEEP_DBG=1 timeout 300 $B t12,4,2,6144,6144,6250,19,4 t24,3,2,6144,6144,6250,19,2 t31,3,2,6144,6144,6250,19,1 t31,2,2,6144,6144,6250,19,2 > gpurun_out/bs_c.log 2>&1; echo "c rc=$?"
cat gpurun_out/bs_a.log gpurun_out/bs_b.log gpurun_out/bs_c.log | grep -v "sweep [12]: VB"
