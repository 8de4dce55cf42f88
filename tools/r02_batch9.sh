#!/bin/bash
# sharded persistent kernel: folded exchange on/off at N GPUs (weak scaling headline only)
out=gpurun_out; mkdir -p $out
N=$1
for fold in 1 0; do
IPXGPU_XFOLD=$fold timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$fold bench.py --gpus $N --steps 10 --warmup 3 --no-c5 > $out/r02l_n${N}_fold$fold.json 2> $out/r02l_n${N}_fold$fold.err; echo "N=$N fold=$fold rc=$?"
python - <<P
import json
lines=[l for l in open("$out/r02l_n${N}_fold$fold.json") if l.startswith("{")]
d=json.loads(lines[-1])
print("N=$N fold=$fold value=%.0f apply_us=%.1f parity=%s bitident=%s" % (d["value"], d["roofline"]["apply_us_in_loop"], d["parity"]["ok"], d["parity"]["ranks_bit_identical"]))
P
tail -2 $out/r02l_n${N}_fold$fold.err | cut -c1-300
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "sharded" 2>&1 | tail -3
