B=./ipx_b200/_build/bsweep_bench
for extra in 0 16384 32768; do echo "=== extra smem $extra"; BSWEEP_EXTRA_SMEM=$extra timeout 300 $B 31,4,2,8192,8192,6250,14,0 2>&1 | grep -E "APPLY|flushed"; done
for vb in 4096 6144; do echo "=== VB $vb"; timeout 300 $B 31,4,2,$vb,$vb,6250,$((114688/vb)),0 2>&1 | grep -E "sweep [12]: VB|APPLY|flushed"; done
echo "=== NBUF 3 VB 4096"; timeout 300 $B 31,4,3,4096,4096,6250,28,0 2>&1 | grep -E "sweep [12]: VB|APPLY|flushed"
echo "=== SB2 11112 K2 8"; timeout 300 $B 31,4,2,8192,8192,11112,8,0 2>&1 | grep -E "sweep [12]: VB|APPLY|flushed"
