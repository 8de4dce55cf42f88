IPXGPU_FUSED_TRACE=2 timeout 300 python tools/profile_apply.py --reps 5 --pcr 6 2>&1 | tail -11
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused5.json 2> gpurun_out/bench_fused5.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_fused5.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "apply_us", d["roofline"]["apply_us_in_loop"])
PY
tail -5 gpurun_out/bench_fused5.err
