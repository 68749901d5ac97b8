# round 2, session 2, call 4: tail imbalance of the event kernel -- per-warp loop-exit clocks and
# dynamic chunk scheduling variants of the product kernel
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2_k1_r3b.txt
for v in base clock dyn1 dyn4 dyn16 dyn4c base dyn4; do
  timeout 120 scripts/exp/k1_r3_$v >> $O/r2_k1_r3b.txt 2>&1
done
for v in base20 dyn4_20 base4 dyn4_4; do
  timeout 120 scripts/exp/k1_r3_$v 50000000 >> $O/r2_k1_r3b.txt 2>&1
done
timeout 120 scripts/exp/k1_r3_base 1000000 >> $O/r2_k1_r3b.txt 2>&1
timeout 120 scripts/exp/k1_r3_dyn4 1000000 >> $O/r2_k1_r3b.txt 2>&1
timeout 120 scripts/exp/k1_r3_base 10000000 >> $O/r2_k1_r3b.txt 2>&1
timeout 120 scripts/exp/k1_r3_dyn4 10000000 >> $O/r2_k1_r3b.txt 2>&1
cat $O/r2_k1_r3b.txt
