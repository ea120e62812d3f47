"""GEMM shape sweep on the GPU: microseconds and TFLOP/s per kernel path
(1 = CUDA-core SGEMM/GEMV, 2 = tcgen05 split-bf16) for the shapes the hot path
actually launches.  Usage: python tools/gemm_sweep.py [--ncu]  (under gpurun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.package()
import importlib

ck = importlib.import_module("pockettts_b200.checkpoint")
rt = pkg.NativeCudaRuntime(tensors=ck.make_checkpoint(ck.ArchConfig.small(), seed=7), max_batch=8,
                           max_seq_tokens=128, max_frames=16, mimi_group_frames=16)
SHAPES = [  # (name, M, K, N, lda)
    ("step qkv B=256", 256, 1024, 3072, 0), ("step out B=256", 256, 1024, 1024, 0),
    ("step l1 B=256", 256, 1024, 4096, 0), ("step l2 B=256", 256, 4096, 1024, 0),
    ("step qkv B=64", 64, 1024, 3072, 0), ("step l2 B=64", 64, 4096, 1024, 0),
    ("step qkv B=16", 16, 1024, 3072, 0), ("flow mlp B=256", 256, 512, 512, 0),
    ("flow ada B=256", 256, 512, 10240, 0),
    ("mimi qkv 8k rows", 8192, 512, 1536, 0), ("mimi ff1", 8192, 512, 2048, 0), ("mimi ff2", 8192, 2048, 512, 0),
    ("initconv", 8192, 3584, 512, 512), ("up1", 8192, 1024, 1536, 512), ("res1.c1", 49152, 768, 128, 256),
    ("res1.c2", 49152, 128, 256, 0), ("up2", 49152, 512, 640, 256), ("res2.c1", 245760, 384, 64, 128),
    ("res2.c2", 245760, 64, 128, 0), ("up3", 245760, 256, 256, 128), ("res3.c1", 983040, 192, 32, 64),
    ("res3.c2", 983040, 32, 64, 0), ("big square", 8192, 4096, 4096, 0),
]
only = sys.argv[1:] and [s for s in SHAPES if sys.argv[1] in s[0]] or SHAPES
print(f"{'shape':20s} {'M':>7s} {'K':>5s} {'N':>6s} | {'cuda-core us':>12s} {'TF/s':>7s} | {'tcgen05 us':>11s} {'TF/s':>7s}")
for name, M, K, N, lda in only:
    fl = 2.0 * M * K * N
    reps = 5 if fl > 5e10 else 20
    a = rt.gemm_bench(M, K, N, path=1, reps=reps, lda=lda) if os.environ.get("SWEEP_CUDA_CORE") else float("nan")
    b = rt.gemm_bench(M, K, N, path=int(os.environ.get('GEMM_PATH', '4')), reps=reps, lda=lda)
    print(f"{name:20s} {M:7d} {K:5d} {N:6d} | {a:12.1f} {fl / a / 1e6:7.1f} | {b:11.1f} {fl / b / 1e6:7.1f}")
rt.close()
