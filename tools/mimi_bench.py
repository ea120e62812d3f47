"""Mimi decode micro-benchmark: one ragged batch of latent frames through latent->Mimi projection,
decoder transformer and the SEANet stack, device time and per-class / per-shape split.
Usage (under gpurun): python tools/mimi_bench.py [frames_per_seq n_seqs]"""
import os, sys, importlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.package()
ck = importlib.import_module("pockettts_b200.checkpoint")
import numpy as np
fps = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nseq = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rt = pkg.NativeCudaRuntime(tensors=ck.make_checkpoint(ck.ArchConfig.full()), max_batch=8, max_seq_tokens=128,
                           max_frames=16, mimi_group_frames=fps * nseq)
lat = np.random.default_rng(3).standard_normal((fps * nseq, 32)).astype(np.float32)
sizes = [fps] * nseq
for _ in range(3):
    rt.decode_latents(lat, sizes)
rt.stage_times(reset=True)
n = 5
for _ in range(n):
    rt.decode_latents(lat, sizes)
st = rt.stage_times()
fr = fps * nseq
print(f"mimi decode of {nseq} x {fps} frames: {st['mimi_ms'] / n:.3f} ms device -> {fr * 0.08 / (st['mimi_ms'] / n / 1e3):.0f} audio-s/s, "
      f"{1e3 * st['mimi_ms'] / n / fr:.2f} us/frame", flush=True)
rt.profile(True)
rt.decode_latents(lat, sizes)
p = rt.profile_report()
rt.profile(False)
for k, v in sorted(p.items(), key=lambda kv: -kv[1]["ms"])[:24]:
    print(f"   {k:40s} {v['launches']:3d} x {1e3 * v['ms'] / v['launches']:8.1f} us  {v['flops'] / max(v['ms'], 1e-9) / 1e9:7.1f} TF/s "
          f"{v['bytes'] / max(v['ms'], 1e-9) / 1e6:7.0f} GB/s")
rt.close()
