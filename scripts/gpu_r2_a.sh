# round 2, first GPU pass: parity tests, default bench line, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench exit $?"
tail -c 3000 gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
