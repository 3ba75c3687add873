# what the round-2 8-GPU numbers in profiles/ came from: sharded tests, then config 3 on 8 B200s (run under gpurun --gpus 8)
summ() { python - "$1" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print(d["n_gpus"], round(d["ms_per_step"],1), round(d["value"],2), (d.get("e2e") or {}).get("ms_per_step"), d.get("per_step"), d.get("alloc_host_ms"))
        print("  ", {k: round(v,1) for k,v in d["stage_ms"].items()})
        print("  ", {k: v["ms"] for k,v in r["per_kernel"].items()}, r.get("sharded_graph_phases_ms"))
        print("  ", d.get("result_digest"))
P
}
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --config c3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2_c3_final.log 2> gpurun_out/r2_c3_final.err
summ gpurun_out/r2_c3_final.log
