# round 2, session 3: VEGAS+ block sizes in the product: plus tests + c3 bench
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "plus or Plus" > $O/r2z8_tests.log 2>&1
echo "pytest exit $?" >> $O/r2z8_tests.log
tail -3 $O/r2z8_tests.log
for i in 1 2; do
timeout 600 python bench.py --workload c3 --no-cpu-baseline --no-table > $O/r2z8_bench_c3.json 2> $O/r2z8_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z8_bench_c3.json').read().strip().splitlines()[-1])
r=d['roofline']
print('c3', d['value'], d['ms_per_step'], r['frac'], r['kernel_ms'], d['e2e']['value'])
PY
done
