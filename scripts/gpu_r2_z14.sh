mkdir -p gpurun_out
O=gpurun_out
: > $O/r2z_d20.txt
for rep in 1 2; do for f in scripts/exp/k1_r3_o20*; do timeout 120 $f 40000000 >> $O/r2z_d20.txt 2>&1; done; done
sort $O/r2z_d20.txt | cut -c1-130
