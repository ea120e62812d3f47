# Batch-1 frame kernel: exchange variant x L2 look-ahead depth (run under gpurun).
cd $GRAFT_REPO_ROOT
O=gpurun_out
for ll in 1 0; do
  echo "== pytest LL=$ll AHEAD=2"
  PTTS_FUSED_LL=$ll PTTS_FUSED_L2_AHEAD=2 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
done
: > $O/fused_matrix.txt
for ll in 0 1; do for a in 0 1 2 3 4; do
  echo -n "LL=$ll L2_AHEAD=$a  " | tee -a $O/fused_matrix.txt
  PTTS_FUSED_LL=$ll PTTS_FUSED_L2_AHEAD=$a timeout 300 python tools/step_bench.py 1 2>&1 | head -1 | tee -a $O/fused_matrix.txt
done; done
for ll in 0 1; do
  PTTS_FUSED_LL=$ll PTTS_FUSED_L2_AHEAD=2 PTTS_FUSED_TIMELINE=1 timeout 300 python tools/step_bench.py 1 2>&1 | grep timeline > $O/timeline_ll${ll}_a2.txt
done
cut -c1-1200 $O/timeline_ll1_a2.txt
