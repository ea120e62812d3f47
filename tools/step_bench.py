"""AR-step micro-benchmark: B live sequences (45-token prompts), CUDA-graph step time on the
device, then one profiled round (events around every launch, graphs off) with the per-class
and per-shape split.  Usage (under gpurun): python tools/step_bench.py [B ...]"""
import os, sys, importlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.package()
ck = importlib.import_module("pockettts_b200.checkpoint")
wl = importlib.import_module("pockettts_b200.workload")
import numpy as np
Bs = [int(a) for a in sys.argv[1:]] or [1, 256]
tensors = ck.make_checkpoint(ck.ArchConfig.full())
if os.environ.get("STEP_BENCH_BF16_EXACT"):  # a checkpoint stored as BF16: the two-product GEMM instantiations
    tensors = ck.to_bf16_exact(tensors)
rt = pkg.NativeCudaRuntime(tensors=tensors, max_batch=max(Bs), max_seq_tokens=512,
                           max_frames=256, mimi_group_frames=64)
c2 = wl.paragraph_200_words()[0]
rng = np.random.default_rng(1)
for B in Bs:
    seqs = [rt.seq_open(c2.tokens, reserve_frames=200) for _ in range(B)]
    noise = (rng.standard_normal((200, B, 32)) * 0.7).astype(np.float32)
    for i in range(6):
        rt.step(seqs, noise[i])
    rt.stage_times(reset=True)
    n = 48
    t0 = time.perf_counter()
    for i in range(n):
        rt.step(seqs, noise[6 + i])
    wall = (time.perf_counter() - t0) / n
    st = rt.stage_times()
    print(f"B={B}: step {1e3 * st['step_ms'] / n:.1f} us device, {1e6 * wall:.1f} us wall, "
          f"{st['kernel_launches'] / n:.0f} launches/step, kv_len~{c2.tokens.size + 30}", flush=True)
    rt.profile(True)
    for i in range(8):
        rt.step(seqs, noise[60 + i])
    p = rt.profile_report()
    rt.profile(False)
    tot = sum(v["ms"] for k, v in p.items() if not k.startswith("shape/"))
    print(f"   profiled (graphs off): sum of kernels {1e3 * tot / 8:.1f} us/step")
    for k, v in sorted(p.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = v["bytes"] / (v["ms"] * 1e6) if v["ms"] else 0
        print(f"   {k:36s} {v['launches'] / 8:5.1f}/step {1e3 * v['ms'] / v['launches']:7.1f} us each "
              f"{1e3 * v['ms'] / 8:7.1f} us/step {gbs:7.0f} GB/s")
    for s in seqs:
        s.close()
rt.close()
