# compute-sanitizer memcheck over the kernels that changed in the last session of round 2 (K5 / K6 / multi-banding:
# row-blocked node and edge tables, per-lane bulk copies)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest -q -x -m gpu \
  tests/test_gpu_reduced.py tests/test_gpu_reduced_cal.py tests/test_gpu_multiband.py tests/test_gpu_edge_cases.py \
  > gpurun_out/r2b_sanitizer.log 2>&1
echo "exit $?" >> gpurun_out/r2b_sanitizer.log
grep "ERROR SUMMARY\|passed\|failed\|exit" gpurun_out/r2b_sanitizer.log | sort | uniq -c | tail -8
