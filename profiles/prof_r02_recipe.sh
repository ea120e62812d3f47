# Round-2 profiling recipe (run under gpurun, one GPU): ncu launch list of the reduced bench command and
# `ncu --set full` of the dominant kernel (tc_gemm_kernel on the AR-step shapes) and of the fused decode
# attention (fp16 cache).  Every command runs plain first; reports are exported on the box.
cd $GRAFT_REPO_ROOT
O=gpurun_out
A="python bench.py --utterances 128 --steps 1 --warmup 1 --no-cpu-baseline --no-latency"
$A > $O/plainA.log 2> $O/plainA.err && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 8000 --csv --log-file $O/launches_r02.csv $A > $O/ncuA.log 2>&1
B="python tools/step_bench.py 256"
$B > $O/plainB.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 600 -c 8 -f -o /tmp/tc $B > $O/ncu_tc.log 2>&1
ncu -i /tmp/tc.ncu-rep --page raw --csv > $O/tcstep_r02_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_paged_decode -s 60 -c 2 -f -o /tmp/at $B > $O/ncu_at.log 2>&1
ncu -i /tmp/at.ncu-rep --page raw --csv > $O/attn_r02_raw.csv 2>/dev/null
rm -f /tmp/tc.ncu-rep /tmp/at.ncu-rep
du -sh $O
