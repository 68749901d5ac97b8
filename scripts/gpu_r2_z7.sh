# round 2, session 3: VEGAS+ event kernel block sizes at other dimensions (uniform allocations of ~5e7 events)
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2z_plus2.txt
for rep in 1 2; do
for t in 0 640 768; do
  timeout 120 scripts/exp/plus_r3_d4t$t 5000 >> $O/r2z_plus2.txt 2>&1
  timeout 120 scripts/exp/plus_r3_d6t$t 12000 >> $O/r2z_plus2.txt 2>&1
  timeout 120 scripts/exp/plus_r3_d10t$t 48000 >> $O/r2z_plus2.txt 2>&1
  timeout 120 scripts/exp/plus_r3_d12t$t 12000 >> $O/r2z_plus2.txt 2>&1
done
done
sort $O/r2z_plus2.txt
