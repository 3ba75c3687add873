timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --het 100 > gpurun_out/r2_het_n1.log 2> gpurun_out/r2_het_n1.err
python - <<P
import json
for l in open("gpurun_out/r2_het_n1.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d.get("per_step"), {k: round(v,1) for k,v in d["stage_ms"].items()}, d["config"].get("edges"), d.get("result_digest"))
P
