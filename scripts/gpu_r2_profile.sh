# round 2 profiling pass (1 GPU): launch list of the default bench command, full ncu capture of
# the sg8 event kernel, metrics of the tail kernels, compute-sanitizer on the small workloads.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv \
    --log-file gpurun_out/r2_launches_sg8.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-table > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 6 -c 1 -f \
    -o gpurun_out/r2_prof_sg8 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-table > gpurun_out/r2_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:finalize_epilogue -s 6 -c 1 -f \
    -o gpurun_out/r2_prof_epilogue \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-table > gpurun_out/r2_ncu_epi.log 2>&1
ncu --set full --clock-control none -k regex:plus_ -s 6 -c 2 -f \
    -o gpurun_out/r2_prof_c3 \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-table > gpurun_out/r2_ncu_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
