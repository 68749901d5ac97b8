mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "symgauss or c1 or golden" > $O/r2z12_tests.log 2>&1; tail -2 $O/r2z12_tests.log
for i in 1 2 3; do
timeout 600 python bench.py --workload c1 --no-cpu-baseline --no-table > $O/r2z12_bench_c1.json 2> $O/r2z12_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z12_bench_c1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('c1', d['value'], d['ms_per_step'], r['frac'], r['kernel_ms'], r['epilogue_kernel_ms'], d['e2e']['value'])
PY
done
