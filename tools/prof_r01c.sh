# Round-1c profiling recipe (run under gpurun): `ncu --set full` of the kernels added late in
# the round, exported to small CSV / text summaries on the box.
cd $GRAFT_REPO_ROOT
O=gpurun_out
cap() { # name, kernel regex, skip, count, command...
  local name=$1 k=$2 s=$3 c=$4; shift 4
  "$@" > $O/plain_$name.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o /tmp/$name "$@" > $O/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > /tmp/${name}_src.csv 2>/dev/null
  python tools/ncu_hot.py /tmp/${name}_src.csv 30 > $O/${name}_hot.txt 2>&1
  rm -f /tmp/$name.ncu-rep /tmp/${name}_src.csv
}
cap attnpaged attention_paged 1700 2 python tools/step_bench.py 256
cap fused step_fused 12 1 python tools/step_bench.py 1
cap attnmma attention_window_mma 2 1 python tools/mimi_bench.py 64 16
cap seanet "tc_gemm_kernel<64" 6 3 python tools/mimi_bench.py 64 16
du -sh $O
