set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_final_pytest_gpu.txt; cat gpurun_out/r2_final_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; tail -c 600 gpurun_out/r2_final_bench_n1.err
W2RAP_NO_SLAB=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/r2_final_memcheck.log
