#!/bin/bash
# Round 2, batch 5 (2 GPUs): the one-process group behind LpSolver, sharded tests, bench at N=2.
out=gpurun_out
mkdir -p $out
nvidia-smi -L
timeout 120 ipx_b200/_build/lds_bench > $out/r02e_lds_bench.log 2>&1; cat $out/r02e_lds_bench.log
for g in 1 2; do
  IPXGPU_NGPUS=$g timeout 300 python tools/solve_lp.py random:50000:500000:10 --impl gpu --crossover 0 --stop-at-switch -1 --per-iter --out $out/r02e_group_g$g.json > $out/r02e_group_g$g.log 2>&1; echo "group g=$g rc=$?"; grep "^gpu" $out/r02e_group_g$g.log | cut -c1-700; tail -3 $out/r02e_group_g$g.log | cut -c1-300
done
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -x -q -k "group or two_gpus or sharded or kktdiag or kkt_solver" > $out/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/r02e_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $out/r02e_bench_n2.json 2> $out/r02e_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-5000 $out/r02e_bench_n2.json; tail -5 $out/r02e_bench_n2.err
