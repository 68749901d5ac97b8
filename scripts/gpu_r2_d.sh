# round 2, session 2: GPU tests, the driver's two bench commands, and the profiling pass whose
# summaries go to profiles/ (launch list, ncu --set full of the sg8 event kernel, the tail kernel,
# the VEGAS+ kernels), exported to text on the box.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/r2d_smi.txt

timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2d_tests.log 2>&1
echo "pytest exit $?" >> $O/r2d_tests.log
tail -5 $O/r2d_tests.log

timeout 600 python bench.py > $O/r2_bench_default_1gpu.json 2> $O/r2d_bench.err
tail -c 600 $O/r2_bench_default_1gpu.json
timeout 600 python bench.py --impl reference > $O/r2_bench_reference_1gpu.json 2>> $O/r2d_bench.err
tail -c 400 $O/r2_bench_reference_1gpu.json

# launch list of the default bench command (short: the share of the step is what matters)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv \
    --log-file $O/r2_launches_sg8.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-table > $O/r2_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 6 -c 1 -f \
    -o $O/r2_prof_sg8 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-table > $O/r2_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:finalize_epilogue -s 6 -c 1 -f \
    -o $O/r2_prof_epilogue \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-table > $O/r2_ncu_epi.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:plus_ -s 6 -c 2 -f \
    -o $O/r2_prof_c3 \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-table > $O/r2_ncu_c3.log 2>&1
for r in r2_prof_sg8 r2_prof_epilogue r2_prof_c3; do
  ncu -i $O/$r.ncu-rep --page details > $O/${r}_details.txt 2>&1
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>&1
done
ncu -i $O/r2_prof_sg8.ncu-rep --page source --csv > $O/r2_prof_sg8_source.csv 2>&1
ls -la $O/
