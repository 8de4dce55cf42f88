#!/bin/bash
# Round 2, batch 3: cross-tile packed rows.
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r02c_pytest.log
for pack in 1 0; do
IPXGPU_BAND_PACK=$pack timeout 300 python bench.py --no-cpu-baseline --no-c5 --no-ipm > $out/r02c_bench_pack$pack.json 2> $out/r02c_bench_pack$pack.err; echo "bench pack=$pack rc=$?"
python - <<P
import json
d=json.load(open("$out/r02c_bench_pack$pack.json")); r=d["roofline"]
print("pack=$pack value=%.0f apply_us=%.2f frac=%.3f iso=%.1f s1=%.1f s2=%.1f parity=%s" % (d["value"], r["apply_us_in_loop"], r["frac"], r["apply_us_isolated_l2_flushed"], r["sweep1_us"], r["sweep2_us"], d["parity"]["ok"]))
P
done
IPXGPU_FUSED_TRACE=1 timeout 200 python bench.py --no-cpu-baseline --no-c5 --no-ipm --no-parity --steps 1 --warmup 3 2>&1 >/dev/null | tail -1 | cut -c1-500
timeout 600 ipx_b200/_build/bsweep_bench 100000 1000000 10 31,4,2,8192,8192,6250,14,0 31,4,2,8192,4096,12500,14,0 31,4,2,8192,4096,6250,28,0 31,4,3,4096,4096,6250,28,0 31,4,4,4096,4096,6250,28,0 31,4,2,4096,4096,12500,14,0 > $out/r02c_bsweep.log 2>&1; echo "bsweep rc=$?"; grep -E "cfg|sweep [12]:|APPLY|flushed" $out/r02c_bsweep.log | cut -c1-220
