B=./ipx_b200/_build/bsweep_bench
timeout 300 $B 3000 40000 10 31,4,2,1024,2048,700,5,0 2>&1 | grep -E "relerr"
BSWEEP_TRACE=1 timeout 300 $B 31,4,2,8192,8192,6250,14,0 2>&1 | grep -E "cfg|APPLY|relerr|TRACE|min "
