# End-of-round profiling recipe (run under gpurun): ncu launch list of the reduced bench
# command + `ncu --set full` of the fused residual-block kernel, exported on the box.
cd $GRAFT_REPO_ROOT
O=gpurun_out
A="python bench.py --utterances 128 --steps 1 --warmup 1 --no-cpu-baseline --no-latency"
$A > $O/plainA.log 2> $O/plainA.err && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 8000 --csv --log-file $O/launches_r01d.csv $A > $O/ncuA.log 2>&1
B="python tools/mimi_bench.py 64 16"
$B > $O/plain_rb.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:resblock_kernel -s 4 -c 2 -f -o /tmp/rb $B > $O/ncu_rb.log 2>&1
ncu -i /tmp/rb.ncu-rep --page raw --csv > $O/rb_raw.csv 2>/dev/null
ncu -i /tmp/rb.ncu-rep --page source --csv > /tmp/rb_src.csv 2>/dev/null
python tools/ncu_hot.py /tmp/rb_src.csv 40 > $O/rb_hot.txt 2>&1
rm -f /tmp/rb.ncu-rep /tmp/rb_src.csv
du -sh $O
