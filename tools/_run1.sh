B=./ipx_b200/_build/bsweep_bench
BSWEEP_TRACE=1 timeout 300 $B 31,4,2,8192,8192,6250,14,0 31,4,3,5460,5460,6250,20,0 2>&1 | grep -E "cfg|APPLY|TRACE|min "
