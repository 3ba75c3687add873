W2RAP_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config c3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2_c3_trace.log 2> gpurun_out/r2_c3_trace.err
grep -E "^\[w2rap\]" gpurun_out/r2_c3_trace.err | tail -120
python - <<P
import json
for l in open("gpurun_out/r2_c3_trace.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["value"], d.get("per_step"), d.get("alloc_host_ms"), d.get("stage_ms"), d.get("kernel_ms"), d.get("phases"), d.get("result_digest"))
P
