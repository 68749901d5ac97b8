# round 2, session 3: VEGAS+ event kernel standalone (6561 cubes x 7620 events = the c3 shape): block sizes x histogram layout
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2z_plus.txt
for rep in 1 2; do
for v in t0p t0n t896p t896n t768p t768n t640p t640n; do
  timeout 120 scripts/exp/plus_r3_$v >> $O/r2z_plus.txt 2>&1
done
done
cat $O/r2z_plus.txt
