N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_bench_n$N.json 2> gpurun_out/r2_final_bench_n$N.err
python - $N <<'P'
import json,sys
for l in open("gpurun_out/r2_final_bench_n%s.json" % sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print(d["n_gpus"], round(d["ms_per_step"],1), round(d["value"],2), (d.get("e2e") or {}).get("ms_per_step"), d.get("per_step"))
        print("  ", {k: round(v,1) for k,v in d["stage_ms"].items()})
        print("  ", d.get("result_digest"))
P
