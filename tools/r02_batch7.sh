#!/bin/bash
# Round 2, batch 7 (8 GPUs): the driver's scaling line at N = 8 and N = 4; the LpSolver group at G = 8.
out=gpurun_out
mkdir -p $out
for n in 8 4; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 > $out/r02i_bench_n$n.json 2> $out/r02i_bench_n$n.err ) 2>&1 | grep real; echo "bench n=$n rc=$?"
python - <<P
import json
lines=[l for l in open("$out/r02i_bench_n$n.json") if l.startswith("{")]
d=json.loads(lines[-1]); c=d["north_star_c5"]
print("N=$n value=%.0f %s apply_us=%.1f parity=%s | c5 %.0f matvec/s apply_us=%.0f other_us=%.0f parity=%s exchange=%s" % (d["value"], d["unit"], d["roofline"]["apply_us_in_loop"], d["parity"]["ok"], c["cr_matvecs_per_sec"], c["apply_us"], c["other_us_per_iter"], c["parity"]["ok"], c["exchange"][:60]))
P
tail -3 $out/r02i_bench_n$n.err
done
for g in 8 1; do
( time IPXGPU_NGPUS=$g timeout 900 python tools/solve_lp.py random:1000000:20000000:5 --impl gpu --crossover 0 --stop-at-switch -1 --maxiter 4 --per-iter --out $out/r02i_c5_group_g$g.json > $out/r02i_c5_group_g$g.log 2>&1 ) 2>&1 | grep real; echo "c5 group g=$g rc=$?"; grep "^gpu" $out/r02i_c5_group_g$g.log | cut -c1-900; tail -6 $out/r02i_c5_group_g$g.log | cut -c1-200
done
