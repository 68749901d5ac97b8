# round 2, session 3: event-kernel block-size sweep, d = 9 ... 20 (and 2, 6, 10)
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2z_threads2.txt
for rep in 1 2; do
for f in scripts/exp/k1_r3_e*; do
  timeout 120 $f 40000000 >> $O/r2z_threads2.txt 2>&1
done
done
sort -t= -k2 -n $O/r2z_threads2.txt | sort -s -k2,2V | cut -c1-112
