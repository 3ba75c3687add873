# final single-GPU verification of the round: GPU test suite, smoke under racecheck (pool allocator: per-buffer bounds)
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_final_pytest_gpu.txt
W2RAP_NO_SLAB=1 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_racecheck.log 2>&1; echo racecheck rc=$?; tail -4 gpurun_out/r2_final_racecheck.log
