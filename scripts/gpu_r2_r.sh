mkdir -p gpurun_out
O=gpurun_out
: > $O/r2_k1_r3d.txt
for v in base quad base quad; do timeout 120 scripts/exp/k1_r3_$v >> $O/r2_k1_r3d.txt 2>&1; done
for v in base20 quad20 base4 quad4; do timeout 120 scripts/exp/k1_r3_$v 50000000 >> $O/r2_k1_r3d.txt 2>&1; done
cat $O/r2_k1_r3d.txt
