"""Device time of the batch-1 AR frame for each frame-kernel variant (grid barrier / flag-in-data /
weight ring), one engine per variant in one process.  Usage (under gpurun): python tools/frame_b1.py [--frames 200]
With PTTS_FUSED_TIMELINE=1 the engine prints CTA 0's phase timeline of frame 20."""
import argparse
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--variants", default="barrier,ll,ring")
args = ap.parse_args()
pkg = ge.package()
ck = importlib.import_module("pockettts_b200.checkpoint")
import numpy as np

tensors = ck.make_checkpoint(ck.ArchConfig.full())
ENV = {"barrier": {"PTTS_FUSED_LL": "0", "PTTS_FUSED_RING": "0"}, "ll": {"PTTS_FUSED_LL": "1", "PTTS_FUSED_RING": "0"},
       "ring": {"PTTS_FUSED_LL": "0", "PTTS_FUSED_RING": "1"}, "unfused": {"PTTS_NO_FUSED_STEP": "1"}}
toks = np.random.default_rng(1).integers(0, 4000, size=40).astype(np.int64)
noise = pkg.make_noise(args.frames + 40, 0.7, 2)
ref = None
for v in args.variants.split(","):
    for k in ("PTTS_FUSED_LL", "PTTS_FUSED_RING", "PTTS_NO_FUSED_STEP"):
        os.environ.pop(k, None)
    os.environ.update(ENV[v])
    rt = pkg.NativeCudaRuntime(tensors=tensors, max_batch=4, max_seq_tokens=512, max_frames=args.frames + 48,
                               mimi_group_frames=64)
    s = rt.seq_open(toks, reserve_frames=args.frames + 44)
    lat = None
    for i in range(40):
        lat, _ = rt.step([s], noise[i])
    rt.stage_times(reset=True)
    for i in range(args.frames):
        lat, _ = rt.step([s], noise[40 + i])
    st = rt.stage_times()
    if ref is None:
        ref = lat.copy()
    print("%-8s frame %.1f us (device, %d frames, kv %d..%d)  launches/frame %.1f  max|lat - first variant| %.2e" %
          (v, 1e3 * st["step_ms"] / args.frames, args.frames, 80, 80 + args.frames, st["kernel_launches"] / args.frames,
           float(np.abs(lat - ref).max())), flush=True)
    s.close()
    rt.close()
