#!/bin/bash
out=gpurun_out; mkdir -p $out
( time timeout 900 python tools/fullsize.py --basis C3 --scale 1.0 --volume-tol 1e30 --kkt-maxiter 100 --out $out/r02o_c3_full_basis.json > $out/r02o_c3.log 2>&1 ) 2>&1 | grep real; echo "c3 rc=$?"; python -c "
import json; d=json.load(open('$out/r02o_c3_full_basis.json'))[0]; print({k:d[k] for k in d if k.startswith('kkt5') or k in ('split_apply_rel_err','split_apply_gpu_ms','split_apply_ref_ms','kkt_solve_ref_s','kkt_solve_gpu_s','cr_ref_s','cr_gpu_s')})"
( time timeout 1500 python -m pytest tests -m gpu -x -q > $out/r02o_pytest.log 2>&1 ) 2>&1 | grep real; tail -3 $out/r02o_pytest.log
timeout 600 python bench.py > $out/r02o_bench.json 2> $out/r02o_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/r02o_bench.json')); r=d['roofline']; print('value=%.0f e2e=%.0f apply_us=%.2f frac=%.3f parity=%s c5=%.0f e2e_ipm=%.2fs launches=%d' % (d['value'], d['e2e']['value'], r['apply_us_in_loop'], r['frac'], d['parity']['ok'], d['north_star_c5']['cr_matvecs_per_sec'], d['e2e_ipm']['time_total'], d['gpu_launches']))"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 --no-ipm --no-parity"
$B > $out/r02o_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches_bench.csv $B > $out/r02o_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
