"""CTA-0 timeline (globaltimer stamps) of one tcgen05 GEMM launch.
Usage (under gpurun): python tools/gemm_timeline.py M K N bn s lda [...]"""
import os, sys, importlib
os.environ["PTTS_GEMM_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.package()
ck = importlib.import_module("pockettts_b200.checkpoint")
rt = pkg.NativeCudaRuntime(tensors=ck.make_checkpoint(ck.ArchConfig.small(), seed=7), max_batch=8,
                           max_seq_tokens=128, max_frames=16, mimi_group_frames=16)
import itertools
args = [int(a) for a in sys.argv[1:]]
PATH = int(os.environ.get('GEMM_PATH', '4'))
for i in range(0, len(args), 6):
    M, K, N, bn, sp, lda = args[i:i + 6]
    rt.tc_force(bn, sp)
    us = rt.gemm_bench(M, K, N, path=PATH, reps=24, lda=lda)
    print(f"M={M} K={K} N={N} bn={bn} s={sp}: {us:.1f} us avg", flush=True)
rt.close()
