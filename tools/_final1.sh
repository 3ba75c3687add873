set -x
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_final_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_good_len|k_minimizer_map|k_scatter_records|k_count_smem|k_insert_solid|k_adjacency|k_links|k_splitter_walk|k_splitter_finish|k_emit_edges|k_path_reads' -c 13 -f -o gpurun_out/r2_final_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_final_full_bench.log 2>&1
ls -la gpurun_out/r2_final_full.ncu-rep gpurun_out/r2_final_launches.csv
