set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -3 gpurun_out/bench_r01b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r01b.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "apply_us", d["roofline"]["apply_us_in_loop"], d.get("cpu_baseline"))
PY
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python tools/profile_apply.py --reps 2 --pcr 6 > gpurun_out/plain_p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'band_sweep|pcr_fused|band_combine' -c 8 -o gpurun_out/r01b_full python tools/profile_apply.py --reps 2 --pcr 6 > gpurun_out/ncu_p.log 2>&1
tail -3 gpurun_out/ncu_b.log gpurun_out/ncu_p.log
ls -la gpurun_out/
