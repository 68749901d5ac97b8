mkdir -p gpurun_out
for v in dy640 dy768 dy1024 st512 st640 st768 dy768 dy1024 st640 st768; do timeout 120 scripts/exp/k1_r3_$v 50000000; done > gpurun_out/r2_me_threads2.txt 2>&1
cat gpurun_out/r2_me_threads2.txt
