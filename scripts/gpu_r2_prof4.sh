# round 2, session 3: ncu --set full of the final VEGAS+ event kernel (c3 shape) and of the d = 20 event kernel (c5)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:plus_event_kernel -s 2 -c 1 -f -o $O/r2s3_prof_c3 scripts/exp/plus_r3_prod > $O/r2s3_ncu_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 2 -c 1 -f -o $O/r2s3_prof_c5 scripts/exp/k1_r3_d20prod 125000000 > $O/r2s3_ncu_c5.log 2>&1
for r in r2s3_prof_c3 r2s3_prof_c5; do
  ncu -i $O/$r.ncu-rep --page details > $O/${r}_details.txt 2>&1
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>&1
  rm -f $O/$r.ncu-rep
  grep -n "Duration\|Issue Slots Busy\|Registers Per Thread" $O/${r}_details.txt | head -4
done
tail -n 2 $O/r2s3_ncu_c3.log; tail -n 2 $O/r2s3_ncu_c5.log
