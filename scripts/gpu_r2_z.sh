# round 2, session 3, call 3: table row address as a C expression vs the inline mad.lo
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2z_addr.txt
for rep in 1 2; do
for v in sg8_mad sg8_lea sg4_mad sg4_lea sg20_mad sg20_lea; do
  timeout 120 scripts/exp/k1_r3_$v >> $O/r2z_addr.txt 2>&1
done
done
cat $O/r2z_addr.txt
