# Round-1 profiling recipe (run under gpurun): ncu launch list of a reduced bench command,
# then `ncu --set full` captures of the top kernels, exported to small CSVs on the box
# (the .ncu-rep files are too large to bring back).
cd $GRAFT_REPO_ROOT
O=gpurun_out
cap() { # name, kernel regex, skip, count, command...
  local name=$1 k=$2 s=$3 c=$4; shift 4
  "$@" > $O/plain_$name.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o /tmp/$name "$@" > $O/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > /tmp/${name}_src.csv 2>/dev/null
  python tools/ncu_hot.py /tmp/${name}_src.csv 40 > $O/${name}_hot.txt 2>&1
  rm -f /tmp/$name.ncu-rep /tmp/${name}_src.csv
}
A="python bench.py --utterances 128 --steps 1 --warmup 1 --no-cpu-baseline --no-latency"
$A > $O/plainA.log 2> $O/plainA.err && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 9000 --csv --log-file $O/launches_r01b.csv $A > $O/ncuA.log 2>&1
cap attn attention_paged 30 1 python tools/step_bench.py 256
cap tcstep tc_gemm 200 4 python tools/step_bench.py 256
cap tcbig tc_gemm 5 1 python tools/gemm_sweep.py big
cap tcinit tc_gemm 5 1 python tools/gemm_sweep.py initconv
cap tcup3 tc_gemm 5 1 python tools/gemm_sweep.py up3
du -sh $O; ls -la $O | tail -24
