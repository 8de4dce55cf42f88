B=./ipx_b200/_build/bsweep_bench
for pin in 0 0.15 0.25 0.35 0.45; do echo "=== pin $pin"; BSWEEP_PIN=$pin timeout 300 $B 31,4,2,8192,8192,6250,14,0 2>&1 | grep -E "APPLY"; done
echo "=== LD=3 (plain ld, allocating) pin n/a"; timeout 300 $B 31,4,2,8192,8192,6250,14,3 2>&1 | grep -E "APPLY"
