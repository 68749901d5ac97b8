# round 2: 2-GPU tests (sharding, timeout/poison), compute-sanitizer on the small kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_api_gpu.py -m gpu -x -q -k "two_gpu" > gpurun_out/r2c_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c_tests.log
tail -12 gpurun_out/r2c_tests.log
export CUDA_VISIBLE_DEVICES=0
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -m gpu -q -x \
   -k "golden or refine or plus_kernel or run_integration_reproduces or vegasflowplus_reproduces or iteration_epilogue or accumulate" > gpurun_out/r2c_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2c_memcheck.log
tail -6 gpurun_out/r2c_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -m gpu -q -x \
   -k "refine or iteration_epilogue or plus_kernel_against_golden or c1-symgauss or plus3a" > gpurun_out/r2c_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r2c_racecheck.log
tail -6 gpurun_out/r2c_racecheck.log
