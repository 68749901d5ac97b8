# round 2, session 3, call 7: pair histograms in the product: GPU tests + the default bench line
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2z5_tests.log 2>&1
echo "pytest exit $?" >> $O/r2z5_tests.log
tail -4 $O/r2z5_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/r2z5_bench.json 2> $O/r2z5_bench.err
tail -3 $O/r2z5_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z5_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('sg8', d['value'], d['ms_per_step'], r['frac'], r['kernel_ms'], 'e2e', d['e2e']['value'])
for k,v in d['workloads'].items(): print(k, '%.4g'%v['value'], '%.4g'%v['ms_per_step'], '%.3f'%v['frac'], '%.4g'%v['kernel_ms'], v.get('frac_implemented_chain'))
PY
