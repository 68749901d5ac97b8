mkdir -p gpurun_out
O=gpurun_out
: > $O/r2_k1_r3c.txt
for v in base faddr base faddr; do timeout 120 scripts/exp/k1_r3_$v >> $O/r2_k1_r3c.txt 2>&1; done
cat $O/r2_k1_r3c.txt
