# round 2, session 2, call 11: driver commands on the final tree (tests, smoke, bench both arms)
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2n_tests.log 2>&1
echo "pytest exit $?" >> $O/r2n_tests.log
tail -3 $O/r2n_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > $O/r2n_bench.json 2> $O/r2n_bench.err ) 2>&1 | grep real
tail -c 300 $O/r2n_bench.json
( time timeout 900 python bench.py --impl reference > $O/r2n_bench_ref.json 2>> $O/r2n_bench.err ) 2>&1 | grep real
