python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/bench_voc_1gpu.json 2>gpurun_out/bench_voc_1gpu.err; tail -c 600 gpurun_out/bench_voc_1gpu.err
python bench.py --workload city --no-cpu-baseline > gpurun_out/bench_city_1gpu.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
python - <<'PY'
import json
for f in ["bench_voc_1gpu","bench_city_1gpu","bench_reference"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("e2e",{}).get("value"), d.get("roofline"), {k:v.get("ms") for k,v in d.get("kernels",{}).items()}, d.get("clocks"), d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
PY
