# round 2, session 3: final ncu evidence of the event kernel with pair histograms (launch list of the
# default bench command, --set full capture with source), compute-sanitizer on the new kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv \
    --log-file $O/r2s3_launches_sg8.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-table > $O/r2s3_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 6 -c 1 -f \
    -o $O/r2s3_prof_sg8 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-table > $O/r2s3_ncu_full.log 2>&1
ncu -i $O/r2s3_prof_sg8.ncu-rep --page details > $O/r2s3_prof_sg8_details.txt 2>&1
ncu -i $O/r2s3_prof_sg8.ncu-rep --page raw --csv > $O/r2s3_prof_sg8_raw.csv 2>&1
rm -f $O/r2s3_prof_sg8.ncu-rep
export CUDA_VISIBLE_DEVICES=0
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -m gpu -q -x \
   -k "golden or refine or plus_kernel or run_integration_reproduces or vegasflowplus_reproduces or iteration_epilogue or accumulate or same_stream or threshold or zoomed" > $O/r2s3_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> $O/r2s3_sanitizer_memcheck.log
tail -4 $O/r2s3_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -m gpu -q -x \
   -k "refine or iteration_epilogue or plus_kernel_against_golden or c1-symgauss or plus3a or same_stream" > $O/r2s3_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> $O/r2s3_sanitizer_racecheck.log
tail -4 $O/r2s3_sanitizer_racecheck.log
head -3 $O/r2s3_launches_sg8.csv; grep -c event_kernel $O/r2s3_launches_sg8.csv
grep -n "Duration\|Issue Slots Busy\|Executed Ipc Active" $O/r2s3_prof_sg8_details.txt | head
