# round 2, session 2, call 6: tests + bench after the VEGAS+ tail change; final ncu evidence of the
# event kernel (launch list, --set full with source), tail kernel and VEGAS+ kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2i_tests.log 2>&1
echo "pytest exit $?" >> $O/r2i_tests.log
tail -3 $O/r2i_tests.log
timeout 600 python bench.py > $O/r2i_bench.json 2> $O/r2i_bench.err
tail -c 200 $O/r2i_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv \
    --log-file $O/r2_launches_sg8.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-table > $O/r2_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 6 -c 1 -f \
    -o $O/r2_prof_sg8 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-table > $O/r2_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:plus_ -s 6 -c 2 -f \
    -o $O/r2_prof_c3 \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-table > $O/r2_ncu_c3.log 2>&1
for r in r2_prof_sg8 r2_prof_c3; do
  ncu -i $O/$r.ncu-rep --page details > $O/${r}_details.txt 2>&1
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>&1
done
ncu -i $O/r2_prof_sg8.ncu-rep --page source --csv > $O/r2_prof_sg8_source.csv 2>&1
ls -la $O/ | head -40
