N=$1
for peer in 1 0; do
IPXGPU_PEER=$peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_peer${peer}.json 2> gpurun_out/bench_n${N}_peer${peer}.err
tail -2 gpurun_out/bench_n${N}_peer${peer}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n${N}_peer${peer}.json").read().strip().splitlines()[-1])
print("N=$N peer=$peer value", d["value"], "e2e", d["e2e"]["value"], "ms_per_step", d["ms_per_step"], "launches", d["gpu_launches"], "frac", d["roofline"]["frac"])
PY
done
