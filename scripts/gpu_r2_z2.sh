# round 2, session 3, call 4: one 50x50 histogram per pair of dimensions (half the shared-memory updates)
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2z_joint.txt
for rep in 1 2; do
for v in sg8_lea sg8_joint1 sg8_joint2 sg4_lea sg4_joint1 sg4_joint2; do
  timeout 120 scripts/exp/k1_r3_$v >> $O/r2z_joint.txt 2>&1
done
done
cat $O/r2z_joint.txt
