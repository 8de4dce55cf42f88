#!/bin/bash
out=gpurun_out; mkdir -p $out
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 10 --warmup 3 --no-c5 > $out/r02n_n${N}.json 2> $out/r02n_n${N}.err; echo "N=$N rc=$?"
python - <<P
import json
lines=[l for l in open("$out/r02n_n${N}.json") if l.startswith("{")]
d=json.loads(lines[-1])
print("N=$N value=%.0f apply_us=%.1f parity=%s bitident=%s" % (d["value"], d["roofline"]["apply_us_in_loop"], d["parity"]["ok"], d["parity"]["ranks_bit_identical"]))
P
done
IPXGPU_XFOLD=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 --no-c5 --no-parity 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=8 fold=0 value=%.0f apply_us=%.1f' % (d['value'], d['roofline']['apply_us_in_loop']))"
