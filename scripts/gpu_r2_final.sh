# round 2 final verification on one GPU (session 3 state): full GPU suite, smoke, examples, both bench arms
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2s3final_tests.log 2>&1
echo "pytest exit $?" >> $O/r2s3final_tests.log
tail -3 $O/r2s3final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for ex in examples/*.py; do
  timeout 300 python $ex > $O/r2s3final_$(basename $ex .py).log 2>&1
  echo "$ex exit $? : $(tail -1 $O/r2s3final_$(basename $ex .py).log | cut -c1-120)"
done
timeout 900 python bench.py > $O/r2s3final_bench.json 2> $O/r2s3final_bench.err
tail -c 200 $O/r2s3final_bench.json; echo
timeout 900 python bench.py --impl reference > $O/r2s3final_bench_ref.json 2>> $O/r2s3final_bench.err
tail -c 200 $O/r2s3final_bench_ref.json; echo
