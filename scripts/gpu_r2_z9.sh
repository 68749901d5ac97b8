# round 2, session 3: event-kernel block sizes with the pair-histogram layout
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2z_threads.txt
for rep in 1 2; do
for v in e4t768 e4t896 e4t1024 e8t768 e8t896 e8t1024 e10t896 e10t1024 e20t640 e20t768 e20t896; do
  timeout 120 scripts/exp/k1_r3_$v 50000000 >> $O/r2z_threads.txt 2>&1
done
done
sort $O/r2z_threads.txt
