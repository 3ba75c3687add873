for mb in 24 48 80 112 0; do
  if [ $mb -gt 0 ]; then export W2RAP_BLOOM_EXACT_MB=$mb; else unset W2RAP_BLOOM_EXACT_MB; fi
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_bloom_$mb.log 2>/dev/null
  python - $mb <<'P'
import json,sys
for l in open("gpurun_out/r2_bloom_%s.log" % sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], round(d["ms_per_step"],1), d["roofline"]["per_kernel"]["k_path_reads"]["ms"], d["roofline"]["per_kernel"]["k_insert_solid"]["ms"], round(d["stage_ms"]["path_ms"],1), d["result_digest"]["check"])
P
done
