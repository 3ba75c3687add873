timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
W2RAP_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_slab_n1.log 2> gpurun_out/r2_slab_n1.err
grep -E "allocation|error" gpurun_out/r2_slab_n1.err | tail -8
python - <<P
import json
for l in open("gpurun_out/r2_slab_n1.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["e2e"], d.get("per_step"), d.get("alloc_host_ms"), d.get("stage_ms"), d.get("result_digest"))
P
