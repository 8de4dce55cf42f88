#!/bin/bash
out=gpurun_out; mkdir -p $out
N=$1
for ov in 1 0; do
IPXGPU_XOVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$ov bench.py --gpus $N --steps 5 --warmup 3 --workload c5 > $out/r02p_c5_n${N}_ov$ov.json 2> $out/r02p_c5_n${N}_ov$ov.err; echo "N=$N overlap=$ov rc=$?"
python - <<P
import json
lines=[l for l in open("$out/r02p_c5_n${N}_ov$ov.json") if l.startswith("{")]
d=json.loads(lines[-1])
print("N=$N overlap=$ov value=%.0f ms_per_step=%.2f apply_us=%.1f parity=%s bitident=%s" % (d["value"], d["ms_per_step"], d["roofline"]["apply_us_in_loop"], d["parity"]["ok"], d["parity"]["ranks_bit_identical"]))
P
tail -2 $out/r02p_c5_n${N}_ov$ov.err | cut -c1-300
done
