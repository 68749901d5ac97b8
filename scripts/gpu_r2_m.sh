# round 2, session 2, call 10: block-size variants of the matrix-element kernels
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2_me_threads.txt
for v in dy512 dy640 dy768 dy896 dy1024 st384 st448 st512 st640 st768; do
  timeout 120 scripts/exp/k1_r3_$v 50000000 >> $O/r2_me_threads.txt 2>&1
done
cat $O/r2_me_threads.txt
