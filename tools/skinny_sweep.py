"""Tile-width / split-K sweep of the tcgen05 GEMM on the AR-step shapes (M = live sequences):
microseconds per launch for every forced (BN, S), TMA-fed A, weights rotated through > L2 worth
of copies.  Usage (under gpurun): python tools/skinny_sweep.py [M ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.package()
import importlib

ck = importlib.import_module("pockettts_b200.checkpoint")
rt = pkg.NativeCudaRuntime(tensors=ck.make_checkpoint(ck.ArchConfig.small(), seed=7), max_batch=8,
                           max_seq_tokens=128, max_frames=16, mimi_group_frames=16)
SHAPES = [("in_proj", 1024, 3072), ("out_proj", 1024, 1024), ("linear1", 1024, 4096), ("linear2", 4096, 1024),
          ("flow_mlp", 512, 512), ("flow_ada", 512, 10240), ("cond_embed", 1024, 512)]
Ms = [int(a) for a in sys.argv[1:]] or [64, 256, 512, 1024]
for M in Ms:
    for name, K, N in SHAPES:
        rt.tc_force(0, 0)
        auto = rt.gemm_bench(M, K, N, path=4, reps=20)
        res = {}
        for bn in (32, 64, 128, 256):
            if bn // 2 >= N:
                continue
            for s in (1, 2, 3, 4, 6, 8):
                if s > K // 64:
                    continue
                tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
                if tiles * s > 148 * 2:
                    continue
                rt.tc_force(bn, s)
                res[(bn, s)] = rt.gemm_bench(M, K, N, path=4, reps=20)
        rt.tc_force(0, 0)
        best = sorted(res.items(), key=lambda kv: kv[1])[:5] or [((0, 0), auto)]
        flops = 2.0 * M * K * N
        print(f"M={M:5d} {name:10s} K={K} N={N}: auto {auto:6.1f} us | best " +
              "  ".join(f"bn{bn}/S{s} {us:5.1f}us" + ("*" if bn and ((M + 127) // 128) * ((N + bn - 1) // bn) * s <= 148 else "")
                        for (bn, s), us in best) + f" | best {flops / best[0][1] / 1e6:6.1f} TF/s", flush=True)
rt.close()
