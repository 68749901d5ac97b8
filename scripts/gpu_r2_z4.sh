# round 2, session 3, call 6: pair-histogram layouts, second sweep (copies vs table copies; partial pairing at d >= 16)
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2z_pairs2.txt
for v in p8 o8_4_1_16_32 p9 o9_4_1_16_32 o11_0_1_8_32 o11_5_1_8_32 o16_0_1_8_16 o16_3_1_8_16 o19_0_1_4_16 o19_4_1_4_16 o20_0_1_4_16 o20_4_1_4_16 o20_5_1_4_16; do
  for rep in 1 2; do timeout 120 scripts/exp/k1_r3_$v 50000000 >> $O/r2z_pairs2.txt 2>&1; done
done
cat $O/r2z_pairs2.txt
