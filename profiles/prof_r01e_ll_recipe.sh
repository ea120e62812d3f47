# ncu --set full of the flag-in-data frame kernel (PTTS_FUSED_LL=1) and of the default barrier kernel.
cd $GRAFT_REPO_ROOT
O=gpurun_out
B="python tools/step_bench.py 1"
for v in 1 0; do
  k=step_fused_ll; [ $v = 0 ] && k=step_fused_kernel
  PTTS_FUSED_LL=$v $B > $O/plain_ll$v.log 2>&1 && PTTS_FUSED_LL=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o /tmp/ll$v $B > $O/ncu_ll$v.log 2>&1
  ncu -i /tmp/ll$v.ncu-rep --page raw --csv > $O/ll${v}_raw.csv 2>/dev/null
  ncu -i /tmp/ll$v.ncu-rep --page source --csv > /tmp/ll${v}_src.csv 2>/dev/null
  python tools/ncu_hot.py /tmp/ll${v}_src.csv 40 > $O/ll${v}_hot.txt 2>&1
  rm -f /tmp/ll$v.ncu-rep /tmp/ll${v}_src.csv
done
du -sh $O
