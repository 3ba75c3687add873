# final single-GPU verification of the round: GPU test suite; one bench run with the step-3 places inside the step
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_final_pytest_gpu.txt
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --places 200 > gpurun_out/r2_bench_c2_places.json 2>/dev/null
python - <<'P'
import json
for l in open("gpurun_out/r2_bench_c2_places.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["places"], d["result_digest"]["check"])
P
