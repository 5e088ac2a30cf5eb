# compute-sanitizer memcheck over the kernels that changed in round 2 (targeted: the whole suite takes too long under it)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest -q -x \
  tests/test_gpu_sampling_front_end.py tests/test_gpu_math_layer.py \
  "tests/test_gpu_parity.py::test_plain_likelihood_and_inner_products_vs_reference" \
  "tests/test_gpu_parity.py::test_calibration_spline_plain_vs_reference" \
  "tests/test_gpu_parity.py::test_time_marginalisation_two_kernel_path_equals_fused_and_oracle" \
  "tests/test_gpu_calmarg.py::test_time_plus_calibration_vs_reference" \
  "tests/test_gpu_reduced.py::test_roq_multibanded_basis_vs_reference" \
  "tests/test_gpu_reduced.py::test_roq_vs_reference" \
  tests/test_gpu_bns.py tests/test_gpu_edge_cases.py > gpurun_out/r2_sanitizer.log 2>&1
echo "exit $?" >> gpurun_out/r2_sanitizer.log
grep -c "ERROR SUMMARY: 0 errors" gpurun_out/r2_sanitizer.log; grep "ERROR SUMMARY\|passed\|failed\|exit" gpurun_out/r2_sanitizer.log | sort | uniq -c | tail -8
