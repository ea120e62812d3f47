"""Tile-width / split-K sweep of the tcgen05 GEMM on the AR-step (skinny M) shapes, with the
weight rotated over >200 MB of copies so nothing is served from L2 (as in the real step).
Usage (under gpurun): python tools/skinny_sweep.py [M ...]"""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.package()
ck = importlib.import_module("pockettts_b200.checkpoint")
rt = pkg.NativeCudaRuntime(tensors=ck.make_checkpoint(ck.ArchConfig.small(), seed=7), max_batch=8,
                           max_seq_tokens=128, max_frames=16, mimi_group_frames=16)
SHAPES = [("qkv", 1024, 3072), ("out", 1024, 1024), ("l1", 1024, 4096), ("l2", 4096, 1024),
          ("flow", 512, 512), ("ada", 512, 10240), ("cond", 1024, 512)]
Ms = [int(a) for a in sys.argv[1:]] or [32, 128, 256, 512]
for M in Ms:
    for name, K, N in SHAPES:
        res = []
        rt.tc_force(0, 0)
        auto = rt.gemm_bench(M, K, N, path=4, reps=24)
        for bn in (32, 64, 128, 256):
            for s in (1, 2, 3, 4, 6, 8):
                if s > K // 64:
                    continue
                rt.tc_force(bn, s)
                res.append((rt.gemm_bench(M, K, N, path=4, reps=24), bn, s))
        res.sort()
        best = " ".join(f"{bn}x{s}:{us:.1f}" for us, bn, s in res[:5])
        print(f"M={M:4d} {name:5s} K={K:5d} N={N:5d} auto={auto:6.1f}us | best {best}", flush=True)
rt.tc_force(0, 0)
rt.close()
