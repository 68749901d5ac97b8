# A/B on one box: c1 (symgauss d=4, 1e6 events/iter) with pair histograms (product) and without
mkdir -p gpurun_out
O=gpurun_out
run() { timeout 600 python bench.py --workload c1 --no-cpu-baseline --no-table > $O/r2z13.json 2> $O/r2z13.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z13.json').read().strip().splitlines()[-1])
r=d['roofline']
print('c1 %.4e step %.3f us K1 %.3f tail %.3f e2e %.4e'%(d['value'], d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['epilogue_kernel_ms']*1e3, d['e2e']['value']))
PY
}
echo "pairs (product)"; run; run
cp vegasflow_b200/lib/libvegasflow_b200.so /tmp/keep.so; cp scripts/exp/libvf_nopairs.so vegasflow_b200/lib/libvegasflow_b200.so
echo "per-dimension histograms"; run; run
cp /tmp/keep.so vegasflow_b200/lib/libvegasflow_b200.so
echo "pairs (product)"; run
