#!/bin/bash
# Round 2, batch 2: dense-column tests, regression, staging experiment, full default bench, IPM per-iteration table.
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_dense_columns.py tests/test_dropin_gpu.py tests/test_gpu_parity.py -x -q > $out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $out/r02b_pytest.log
BSWEEP_DBG=1 timeout 300 ipx_b200/_build/bsweep_bench 100000 1000000 10 31,4,2,8192,8192,6250,14,0 > $out/r02b_bsweep_dbg.log 2>&1; echo "bsweep rc=$?"; grep -E "sweep|APPLY" $out/r02b_bsweep_dbg.log | cut -c1-200
( time timeout 900 python bench.py > $out/r02b_bench.json 2> $out/r02b_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; cut -c1-6000 $out/r02b_bench.json; tail -5 $out/r02b_bench.err
timeout 600 python tools/solve_lp.py random:50000:500000:10 --crossover 0 --stop-at-switch -1 --per-iter --out $out/r02b_periter_50k.json > $out/r02b_periter_50k.log 2>&1; echo "periter rc=$?"; tail -40 $out/r02b_periter_50k.log | cut -c1-400
