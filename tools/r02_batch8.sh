#!/bin/bash
out=gpurun_out
mkdir -p $out
for mode in balance roundrobin; do
IPXGPU_BAND_WARPS=$mode BSWEEP_TRACE=1 timeout 300 ipx_b200/_build/bsweep_bench 100000 1000000 10 31,4,2,8192,8192,6250,14,0 2>&1 | grep -E "sweep [12]:|APPLY|band wait per warp \(mean\)|all warps done|TRACE" | cut -c1-170
IPXGPU_BAND_WARPS=$mode timeout 300 python bench.py --no-cpu-baseline --no-c5 --no-ipm > $out/r02k_bench_$mode.json 2> $out/r02k_bench_$mode.err; echo "bench $mode rc=$?"
python - <<P
import json
d=json.load(open("$out/r02k_bench_$mode.json")); r=d["roofline"]
print("$mode value=%.0f apply_us=%.2f frac=%.3f iso=%.1f s1=%.1f s2=%.1f parity=%s" % (d["value"], r["apply_us_in_loop"], r["frac"], r["apply_us_isolated_l2_flushed"], r["sweep1_us"], r["sweep2_us"], d["parity"]["ok"]))
P
done
