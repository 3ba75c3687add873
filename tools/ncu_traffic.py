#!/usr/bin/env python
"""DRAM traffic per launch of the named kernels from `ncu --set full` reports of the BENCH workload -> profiles/r2_ncu_traffic.json,
which bench.py reads for `roofline.traffic`.  usage: ncu_traffic.py out.json rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import json
import subprocess
import sys

NAMES = {"k_good_len": "k_good_len", "k_minimizer_map": "k_minimizer_map", "k_scatter_records": "k_scatter_records", "k_count_smem": "k_count_smem",
         "k_insert_solid": "k_insert_solid", "k_adjacency": "k_adjacency", "k_links": "k_links", "k_splitter_walk": "k_splitter_walk",
         "k_splitter_finish": "k_splitter_finish", "k_emit_edges": "k_emit_edges", "k_path_reads": "k_path_reads"}


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    res = {}
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}

        def val(r, name):
            v, u = float(r[idx[name]].replace(",", "")), units[idx[name]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        for r in rows[2:]:
            kn = r[idx["Kernel Name"]]
            for key, tag in NAMES.items():
                if kn.startswith(key) or (" " + key) in kn:
                    try:
                        b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
                    except ValueError:
                        continue
                    if b != b:      # nan
                        continue
                    # first launch of each kernel in the capture (the bench runs one read batch resident: one launch per step)
                    res.setdefault(tag, {"dram_bytes_per_launch": b, "duration_ms_under_ncu": float(r[idx["gpu__time_duration.sum"]].replace(",", "")),
                                         "source": "ncu --set full --clock-control none of `python bench.py` (config 2), " + rep.split("/")[-1]})
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
