mkdir -p gpurun_out
for v in st512 st640 st768 st512 st640; do timeout 120 scripts/exp/k1_r3_$v 50000000; done > gpurun_out/r2_st_threads.txt 2>&1
cat gpurun_out/r2_st_threads.txt
