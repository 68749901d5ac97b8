# round 2, session 3, call 5: pair histograms (product form, constexpr chooser) vs per-dimension, d = 2 ... 17
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2z_pairs.txt
for d in 2 3 4 5 6 7 8 9 10 12 13 14 15 16 17; do
  for v in n$d p$d n$d p$d; do timeout 120 scripts/exp/k1_r3_$v 50000000 >> $O/r2z_pairs.txt 2>&1; done
done
cat $O/r2z_pairs.txt
