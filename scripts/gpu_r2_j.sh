# round 2, session 2, call 7: matrix-element rewrite (Drell-Yan half-angle forms) -- parity + bench
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2j_tests.log 2>&1
echo "pytest exit $?" >> $O/r2j_tests.log
tail -15 $O/r2j_tests.log
timeout 600 python bench.py --no-cpu-baseline > $O/r2j_bench.json 2> $O/r2j_bench.err
tail -c 200 $O/r2j_bench.json
