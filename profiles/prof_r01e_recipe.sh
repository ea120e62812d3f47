# Round-1e recipe (run under gpurun): end-of-round bench lines, then `ncu --set full` of the
# flag-in-data batch-1 frame kernel and its CTA-0 timeline.
cd $GRAFT_REPO_ROOT
O=gpurun_out
python bench.py > $O/bench_r01e_default.json 2> $O/bench_r01e_default.err
python bench.py --impl reference > $O/bench_r01e_reference.json 2> $O/bench_r01e_reference.err
B="python tools/step_bench.py 1"
$B > $O/plain_ll.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_fused_ll -s 12 -c 1 -f -o /tmp/ll $B > $O/ncu_ll.log 2>&1
ncu -i /tmp/ll.ncu-rep --page raw --csv > $O/ll_raw.csv 2>/dev/null
ncu -i /tmp/ll.ncu-rep --page source --csv > /tmp/ll_src.csv 2>/dev/null
python tools/ncu_hot.py /tmp/ll_src.csv 40 > $O/ll_hot.txt 2>&1
rm -f /tmp/ll.ncu-rep /tmp/ll_src.csv
PTTS_FUSED_LL=1 PTTS_FUSED_TIMELINE=1 $B 2>&1 | grep timeline > $O/ll_timeline.txt
PTTS_FUSED_LL=0 PTTS_FUSED_TIMELINE=1 $B 2>&1 | grep timeline > $O/barrier_timeline.txt
PTTS_FUSED_LL=0 $B 2>&1 | head -1 > $O/plain_barrier.log
du -sh $O
