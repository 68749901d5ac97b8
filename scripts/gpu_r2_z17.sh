# fixed cost of one event-kernel launch: kernel time vs number of events (148 blocks x 1024 threads = 151552 threads)
mkdir -p gpurun_out
for v in f4 f8 f8p; do for n in 1 151552 303104 606208 1000000 2000000 4000000 10000000 40000000; do scripts/exp/k1_r3_$v $n; done; done 2>&1 | awk '{print $1,$2,$3,$4, $7,$8,$9,$10}' | tee gpurun_out/r2z_fixed.txt
