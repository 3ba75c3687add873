#!/usr/bin/env python
"""bench.py — step-2 K=60 graph build throughput (Gbases/s) on N B200s; see DESIGN.md §Measurement.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--genome-mbp G] [--coverage C] [--impl reference]

One step = one whole pass of step 2 (count -> adjacency -> unipaths -> HBV -> read pathing) over one synthetic read set.
`value` = bases / device time with the read stores already resident in HBM; `e2e` = the same through w2rap_step2_run()
with pinned HOST buffers (H2D of the stores and D2H of graph + paths inside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "step-2 K=60 graph build Gbases/s"


def clocks_sampler(stop, out):
    """SM clock + throttle reasons during the timed region.  In-process NVML (a few microseconds per sample); spawning nvidia-smi
    five times a second re-initialises NVML every time, which takes driver locks and perturbs the very calls being timed."""
    idx = int(os.environ.get("LOCAL_RANK", "0"))
    try:
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(idx)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        while not stop.is_set():
            sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
            r = int(reasons_fn(h))
            act = lambda bit: "Active" if r & bit else "Not Active"
            out.append([str(sm), str(mx), act(0x8), act(0x40), act(0x20), act(0x4)])      # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
            stop.wait(0.05)
        return
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i", str(idx)], capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(1.0)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    reasons = []
    for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
        if any(s[2 + i].lower().startswith("active") for s in samples):
            reasons.append(name)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None, "reasons": reasons}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference_arm(args):
    """The reference's own CPU step 2 (oracle/_ref/w2rap-contigger, all host threads) on a bounded sample of the same workload."""
    import numpy as np
    import w2r_testlib as T
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    genome = int(args.ref_genome_mbp * 1e6)
    rng = np.random.default_rng(1)
    g = T.make_genome(rng, genome, max(1, genome // 50000))
    n_pairs = genome * args.coverage // 500
    rs = T.flatten_reads(*T.simulate_reads(rng, [(g, False, 1.0)], n_pairs, 250))
    kind = "reference" if os.path.exists(T.REF_BIN) else "port"
    times = []
    for it in range(args.warmup + args.steps):
        if kind == "reference":
            d = tempfile.mkdtemp(prefix="w2rap_ref_")
            T.write_fastb_qualp(d, rs)
            _, perf = T.run_reference_step2(d, threads=cores)
            t = perf.get("buildReadQGraph", 0.0) + perf.get("FixPaths", 0.0)
            subprocess.run(["rm", "-rf", d])
        else:
            t0 = time.time()
            T.run_oracle(rs, T.default_params(apply_fixpaths=1))
            t = time.time() - t0
        if it >= args.warmup:
            times.append(t)
    t = sum(times) / len(times)
    v = rs.n_bases / t / 1e9
    sample = "%.1f Mbp random genome + repeats, 2x250 PE at %dx (%d reads, %.0f Mbases); %s" % (
        args.ref_genome_mbp, args.coverage, rs.n, rs.n_bases / 1e6,
        "oracle/_ref/w2rap-contigger --from_step 2 --to_step 2, TIME buildReadQGraph+FixPaths" if kind == "reference" else "oracle/step2_oracle.c")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "bounded sample of the bench workload: " + sample},
            "cpu_baseline": {"value": v, "unit": "Gbases/s", "cores": cores if kind == "reference" else 1, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


NCU_TRAFFIC_RATIO = {"k_minimizer_map<store>": 0.95, "k_count_smem": 0.97}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--genome-mbp", type=float, default=135.0, help="Arabidopsis-sized (BASELINE.json configs[1])")
    ap.add_argument("--coverage", type=int, default=60)
    ap.add_argument("--read-len", type=int, default=250)
    ap.add_argument("--ref-genome-mbp", type=float, default=2.0, help="size of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import w2r_testlib as T
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = T.product_lib()
    if lib.w2rap_step2_device_count() <= local:
        raise SystemExit("no B200 visible: this benchmark has no CPU path")
    err = C.create_string_buffer(1024)
    genome = int(args.genome_mbp * 1e6)
    # strong scaling: ONE read set (a function of the seed only); rank r holds the reads [r*n/N, (r+1)*n/N) of it
    total_reads = (genome * args.coverage // args.read_len) & ~1
    per = (total_reads // world) & ~1
    first = per * rank
    mine = per if rank < world - 1 else total_reads - first
    sp = T.SynthParams(genome, args.read_len, args.coverage, 1000, 0, 0, mine, first)
    h = C.c_void_p()
    comm = C.c_void_p()
    if world > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            if lib.w2rap_step2_comm_unique_id(buf, err, 1024):
                raise SystemExit("nccl id failed: " + err.value.decode())
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        buf = (C.c_uint8 * 128)(*uid.cpu().tolist())
        if lib.w2rap_step2_comm_init(buf, world, rank, local, C.byref(comm), err, 1024):
            raise SystemExit("comm init failed: " + err.value.decode())
    if lib.w2rap_step2_synth(C.byref(sp), local, C.byref(h), err, 1024):
        raise SystemExit("synth failed: " + err.value.decode())
    p = T.default_params(apply_fixpaths=1, device=local)
    hr = T.Reads()
    if lib.w2rap_step2_download_reads(h, C.byref(hr), err, 1024):
        raise SystemExit("download failed: " + err.value.decode())
    n_reads = int(hr.n_reads)
    n_bases = n_reads * args.read_len

    def barrier():
        if dist is not None:
            dist.barrier()

    def step_resident():
        g = T.Graph()
        rc = lib.w2rap_step2_run_sharded_resident(h, C.byref(p), comm, C.byref(g), err, 1024) if world > 1 else lib.w2rap_step2_run_resident(h, C.byref(p), C.byref(g), err, 1024)
        if rc:
            raise SystemExit("run failed: " + err.value.decode())
        t = {n: getattr(g.timings, n) for n, _ in T.Timings._fields_}
        info = (int(g.n_kmer_instances), int(g.n_distinct), int(g.n_solid), int(g.n_edges), int(g.n_edge_bases), int(g.n_pathed), int(g.n_path_edges))
        lib.w2rap_step2_free(C.byref(g))
        return t, info

    def step_e2e():
        g = T.Graph()
        rc = lib.w2rap_step2_run_sharded(C.byref(hr), C.byref(p), comm, C.byref(g), err, 1024) if world > 1 else lib.w2rap_step2_run(C.byref(hr), C.byref(p), C.byref(g), err, 1024)
        if rc:
            raise SystemExit("run failed: " + err.value.decode())
        t = {n: getattr(g.timings, n) for n, _ in T.Timings._fields_}
        d2h = 8 * (int(g.n_edges) + 1) + 4 * int(g.n_edges) * 7 + int(g.n_edge_bases) // 4 + 4 * int(g.n_paths) + 8 * (int(g.n_paths) + 1) + 4 * int(g.n_path_edges)
        lib.w2rap_step2_free(C.byref(g))
        return t, d2h

    for _ in range(args.warmup):
        step_resident()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples), daemon=True)
    th.start()
    barrier()
    tt, info = [], None
    t_wall0 = time.time()
    for _ in range(args.steps):
        t, info2 = step_resident()
        if info is not None and info2 != info:
            raise SystemExit("non-deterministic result between steps: %r vs %r" % (info, info2))
        info = info2
        tt.append(t)
    barrier()
    wall_resident = time.time() - t_wall0
    dev_ms = sum(t["total_ms"] - t["d2h_ms"] for t in tt) / len(tt)      # device time, results left on the device side of the copy
    full_ms = sum(t["total_ms"] for t in tt) / len(tt)
    # end to end through the host-buffer entry point
    step_e2e()
    barrier()
    t0 = time.time()
    e2e_d2h = 0
    e2e_t = []
    for _ in range(args.steps):
        tt_e, e2e_d2h = step_e2e()
        e2e_t.append(tt_e)
    barrier()
    e2e_ms = (time.time() - t0) / args.steps * 1e3
    stop.set()
    th.join(timeout=2)
    total_bases = n_bases
    if dist is not None:
        import torch
        v = torch.tensor([dev_ms, e2e_ms, full_ms], device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, full_ms = [float(x) for x in v.tolist()]
        nb = torch.tensor([n_bases], device="cuda", dtype=torch.int64)
        dist.all_reduce(nb)
        total_bases = int(nb.item())
    if rank != 0:
        if world > 1:
            lib.w2rap_step2_comm_destroy(comm)
            dist.destroy_process_group()
        return
    I, D, S, E, EB, pathed, npe = info
    I = I // world                      # this rank's share of the k-mer instances (the counters are whole-job)
    h2d = int(hr.base_off and 0) + n_reads * ((args.read_len + 3) // 4) + 0
    qbytes = int(T._arr(hr.qual_off, n_reads + 1, "<u8")[-1])
    b_in = n_reads * ((args.read_len + 3) // 4) + qbytes + 12 * n_reads
    b_path = 6 * n_reads + 4 * npe
    alg_bytes_total = 2 * b_in + 34 * I + 48 * S + EB // 2 + b_path
    # dominant kernel: k_minimizer_map, store launch (map: reads the packed bases, writes one record per instance) or k_count_smem
    # (reduce: reads every record back).  SURVEY 8(d) charges 34 B per instance for "written once and read once": 17 B each.
    part_ms = sum(t["count_kernel_ms"] for t in tt) / len(tt)
    region_ms = sum(t["region_ms"] for t in tt) / len(tt)
    b_bases = n_reads * ((args.read_len + 3) // 4) + 14 * n_reads
    if part_ms >= region_ms:
        dom, count_ms, count_launches, alg_bytes_count = "k_minimizer_map<store>", part_ms, tt[-1]["count_launches"], b_bases + 17 * I
    else:
        dom, count_ms, count_launches, alg_bytes_count = "k_count_smem", region_ms, 1, 17 * I      # (+ a handful of fallback launches for oversized partitions)
    peak, peak_src = peaks()
    achieved = alg_bytes_count / (count_ms * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                "--coverage", str(args.coverage), "--ref-genome-mbp", str(args.ref_genome_mbp)], capture_output=True, text=True, timeout=900)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:   # the baseline is reported, never required
            cpu = {"value": None, "unit": "Gbases/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)[:200]}
    line = {
        "metric": METRIC, "value": total_bases / (dev_ms * 1e-3) / 1e9, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "%.0f Mbp synthetic genome + repeat families, 2x%d bp PE at %dx (rank 0 shard: %d reads, %.2f Gbases; whole job = n_gpus shards), min_qual 7, min_freq 4; whole step 2 incl. read pathing + FixPaths" % (
                       args.genome_mbp, args.read_len, args.coverage, n_reads, n_bases / 1e9),
                   "cache": "inputs (%.1f GB) and counting table larger than the 126 MB L2" % (b_in / 1e9),
                   "kmer_instances": I, "distinct": D, "solid": S, "edges": E, "reads_pathed": pathed,
                   "parallelism": "one process per GPU; reads sharded by index, k-mer records routed to owner GPUs by minimiser partition (NCCL all-to-all of exactly sized runs), solid records all-gathered, graph built on every rank, reads pathed by shard"},
        "e2e": {"value": total_bases / (e2e_ms * 1e-3) / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": b_in - 0, "d2h_bytes_per_step": e2e_d2h, "ms_per_step": e2e_ms,
                "inside": {k: sum(t[k] for t in e2e_t) / len(e2e_t) for k in ("h2d_ms", "total_ms", "count_ms", "count_kernel_ms", "region_ms", "path_ms", "d2h_ms", "host_pre_ms", "host_post_ms", "wall_ms")}},
        "gpu_launches": int(tt[-1]["kernel_launches"]) * args.steps,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # ncu --set full of this kernel (profiles/r1_ncu_minimizer_map_count_smem_20mbp.txt, 20 Mbp job): dram read+write =
                     # 0.95x (map) / 0.97x (reduce) the algorithmic bytes; scaled to this workload's bytes per launch
                     "traffic": alg_bytes_count / max(1, count_launches) * NCU_TRAFFIC_RATIO[dom] if dom in NCU_TRAFFIC_RATIO else None,
                     "traffic_source": "ncu dram__bytes_read.sum+dram__bytes_write.sum on the 20 Mbp job, as a ratio to algorithmic bytes, applied to this workload",
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_count / max(1, count_launches), "launches_per_step": count_launches,
                     "kernel_ms_per_step": count_ms, "map_store_ms": part_ms, "reduce_ms": region_ms, "whole_step_algorithmic_GBps": alg_bytes_total / (dev_ms * 1e-3) / 1e9},
        "stage_ms": {k: sum(t[k] for t in tt) / len(tt) for k in ("count_ms", "exchange_ms", "adjacency_ms", "unipath_ms", "hbv_ms", "path_ms", "d2h_ms", "total_ms")},
        "clocks": summarize_clocks(samples),
        "per_step": {"resident_ms": [round(t["total_ms"] - t["d2h_ms"], 2) for t in tt], "resident_count_ms": [round(t["count_ms"], 2) for t in tt],
                     "e2e_wall_ms": [round(t["wall_ms"], 2) for t in e2e_t]},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    lib.w2rap_step2_free_host_reads(C.byref(hr))
    lib.w2rap_step2_release(h)
    if world > 1:
        lib.w2rap_step2_comm_destroy(comm)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
