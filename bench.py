#!/usr/bin/env python
"""bench.py — headline benchmark of the native-cuda PocketTTS backend.

Metric (BASELINE.json): batched audio-seconds per wall-second.  One "step" is
one pass of the whole hot path (prefill -> AR frames + flow head -> Mimi
decode -> PCM) over this rank's shard of the C5 workload ("4096 queued
synthetic utterances sharded across 1/2/4/8 B200"): 4096/8 = 512 utterances of
U{5..60} words per GPU, all submitted at t=0 and continuously batched.  Per-GPU
work is fixed as N grows (weak scaling); at N=8 the job is exactly C5.

  value : throughput with PCM left on the device (inputs are ~24 KB of tokens +
          noise per request, staged at admission)
  e2e   : the same pass through the public C-ABI call with HOST buffers:
          tokens/noise H2D and float PCM D2H inside the timed region
  --impl reference : the reference's CPU path.  No Go toolchain exists in this
          image, so it is the C restatement (oracle/, "port") timed on all
          host cores over a bounded sample of the same workload.

Usage: python bench.py --gpus N --steps K --warmup W [--impl reference]
(for N>1 launch under torch.distributed.run; see the driver contract).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

# stdout carries exactly one JSON line.  Libraries write to file descriptor 1 behind Python's
# back (NCCL prints its version banner there at every debug level above NONE), so fd 1 is pointed
# at stderr for the whole process and the JSON line goes to a private duplicate of the real stdout.
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: str) -> None:
    os.write(_REAL_STDOUT, (line + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

METRIC = "audio-s/wall-s (batched)"
UNIT = "audio-s/s"
TOTAL_UTTERANCES = 4096
MAX_GPUS = 8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native-cuda", choices=["native-cuda", "reference"])
    ap.add_argument("--utterances", type=int, default=TOTAL_UTTERANCES // MAX_GPUS, help="per GPU")
    ap.add_argument("--pool", action="store_true",
                    help="ONE process drives --gpus devices through the in-process pool (ptts_pool_*: least-loaded "
                         "dispatch of one shared queue) instead of one torchrun rank per GPU")
    ap.add_argument("--max-batch", type=int, default=1024)
    ap.add_argument("--mimi-group", type=int, default=2048, help="latent frames per batched Mimi decode workspace")
    ap.add_argument("--arch", default="full", choices=["full", "small"])
    ap.add_argument("--cpu-sample-frames", type=int, default=0,
                    help="cap on frames per chunk of the CPU sample (0 = whole chunks, the default: a cap keeps the "
                         "full prompt prefill and so under-states the CPU path)")
    ap.add_argument("--queue", type=int, default=0,
                    help="strong-scaling mode: this many queued utterances in TOTAL (C5 as written: 4096), split "
                         "over the ranks; default 0 = the driver's weak-scaling contract (--utterances per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


def load_ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel classes captured
    with `ncu --set full`; profiles/ncu_traffic_r0N.json record the commands and the launches averaged."""
    out = {}
    for name in ("ncu_traffic_r01.json", "ncu_traffic_r02.json"):  # later rounds override per class
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                out.update(json.load(f).get("classes", {}))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smmax, reasons = [], None, set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smmax = float(c[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smmax, "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_run(om, chunks, frames_cap, threads):
    """Oracle port on `threads` host threads, one chunk per thread (the reference runs
    N goroutines each on a B=1 loop).  Returns (audio_seconds, wall_seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    frames = [0] * len(chunks)

    def one(i):
        c = chunks[i]
        ms = min(c.max_steps, frames_cap) if frames_cap > 0 else c.max_steps
        force = min(c.force_eos_step, ms - 1 - c.frames_after_eos) if ms > c.frames_after_eos + 1 else 0
        pcm, lat, _ = om.generate_audio(c.tokens, noise=c.noise()[:ms], max_steps=ms, eos_threshold=1e30,
                                        force_eos_step=max(force, 0), frames_after_eos=c.frames_after_eos)
        frames[i] = lat.shape[0]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(len(chunks))))
    dt = time.perf_counter() - t0
    return sum(frames) * 1920 / 24000.0, dt, sum(frames)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and not args.pool:
        # convenience: relaunch under torchrun (the driver does this itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)

    pkg = ge.package()
    import importlib
    ck = importlib.import_module("pockettts_b200.checkpoint")
    wl = importlib.import_module("pockettts_b200.workload")
    sh = importlib.import_module("pockettts_b200.sharding")
    arch = ck.ArchConfig.full() if args.arch == "full" else ck.ArchConfig.small()
    scaling = "weak"
    n_dev = args.gpus if args.pool else 1  # devices THIS process drives
    n_gpus_job = args.gpus if args.pool else world
    if args.pool and args.queue <= 0:
        args.queue = args.utterances * args.gpus  # same total work as the sharded mode
    if args.queue > 0 and args.pool:  # one shared queue, dispatched in-process
        scaling = "strong" if args.queue == TOTAL_UTTERANCES else "weak"
        args.utterances = args.queue
        workload_name = (f"C5 through the in-process pool: {args.queue} queued utterances x U{{5..60}} words, one process, "
                         f"{args.gpus} GPU(s), least-loaded dispatch, <=50-token chunks, forced-EOS schedule, "
                         f"lsd_steps=1, random-init PocketTTS ({args.arch}), seed 20260118")
    elif args.queue > 0:  # C5 as written: one queue of N utterances, whatever the GPU count
        scaling = "strong"
        args.utterances = sh.shard_range(rank, world, args.queue)[1] - sh.shard_range(rank, world, args.queue)[0]
        workload_name = (f"C5: {args.queue} queued utterances in total x U{{5..60}} words split over {world} GPU(s), "
                         f"<=50-token chunks, forced-EOS schedule, lsd_steps=1, random-init PocketTTS ({args.arch}), "
                         f"seed 20260118")
    else:
        workload_name = (f"C5 shard: {args.utterances} utterances/GPU x U{{5..60}} words, <=50-token chunks, "
                         f"forced-EOS schedule, lsd_steps=1, random-init PocketTTS ({args.arch}), seed 20260118")
    DTYPE = "f32 storage/accumulate, split-bf16x3 operands"
    config = {"workload": workload_name,
              "utterances_per_gpu": args.utterances // args.gpus if args.pool else args.utterances,
              "max_batch": args.max_batch,
              "parallelism": (f"in-process pool over {args.gpus} GPU(s), least-loaded dispatch, no collective" if args.pool
                              else "one process per GPU, utterances sharded by rank, no collective"),
              "l2": "working set (weights 341 MB + KV + Mimi workspace, several GB) >> 126 MB L2; no flush needed",
              "weights": "fp32 as stored by the reference", "kv_cache": "fp16 storage, fp32 attention arithmetic",
              "e2e_host_buffers": "tokens / noise from host arrays every step; PCM into caller-owned host buffers "
                                  "allocated once and reused every step (RuntimeGenerateConfig.pcm_out)"}

    # ---------------- reference arm: CPU port on host cores -----------------
    if args.impl == "reference":
        if rank != 0:
            return
        o = ge.oracle()
        cores = os.cpu_count() or 1
        tensors = ck.make_checkpoint(arch)
        om = o.OracleModel(tensors)
        chunks = wl.make_workload(cores, 5, 60, seed=7)[:cores]
        for _ in range(max(args.warmup, 0) and 1):
            cpu_run(om, chunks[: max(1, cores // 2)], 4, cores)
        tot_a = tot_t = 0.0
        for _ in range(args.steps):
            a, t, _ = cpu_run(om, chunks, args.cpu_sample_frames, cores)
            tot_a += a
            tot_t += t
        v = tot_a / tot_t
        cap_txt = f"capped at {args.cpu_sample_frames} frames each" if args.cpu_sample_frames > 0 else "whole chunks"
        sample = f"{cores} chunks (first of the workload), {cap_txt}, one chunk per host thread, per step"
        emit(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1),
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "C restatement of the reference (Go toolchain unavailable); AVX2+FMA dot"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---------------- native-cuda arm ---------------------------------------
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pkg.build_library()
    tensors = ck.make_checkpoint(arch)
    if args.pool:
        rt = pkg.NativeCudaPool(tensors=tensors, devices=list(range(args.gpus)), max_batch=args.max_batch,
                                max_seq_tokens=512, max_frames=256, mimi_group_frames=args.mimi_group)
    else:
        rt = pkg.NativeCudaRuntime(tensors=tensors, device=local_rank, max_batch=args.max_batch, max_seq_tokens=512,
                                   max_frames=256, mimi_group_frames=args.mimi_group)
    first_utt = 0 if args.pool else (sh.shard_range(rank, world, args.queue)[0] if args.queue > 0 else
                                     sh.shard_range(rank, world, args.utterances)[0])
    chunks = wl.make_workload(args.utterances, 5, 60, seed=7, first_utt=first_utt)
    noises = [c.noise() for c in chunks]
    h2d = sum(c.tokens.nbytes + n.nbytes for c, n in zip(chunks, noises))

    # Host PCM buffers of the e2e leg: allocated and touched once, reused by every step, the way a server reuses
    # its output buffers (fresh multi-MB allocations would be page-faulted in by the finisher threads every step).
    outs = [np.empty(c.max_steps * pkg.SAMPLES_PER_FRAME, np.float32) for c in chunks]
    for o in outs:
        o.fill(0.0)  # touch every page now

    def run_pass(no_d2h: bool):
        pend = []
        for c, n, o in zip(chunks, noises, outs):
            cfg = pkg.RuntimeGenerateConfig(max_steps=c.max_steps, eos_threshold=1e30, force_eos_step=c.force_eos_step,
                                            frames_after_eos=c.frames_after_eos, noise=n, no_pcm_d2h=no_d2h,
                                            pcm_out=None if no_d2h else o)
            pend.append(rt.submit(c.tokens, cfg))
        frames, samples = 0, 0
        for p in pend:
            r = rt.wait(p)
            frames += r.n_frames
            samples += r.pcm.size
        return frames, samples

    def timed(no_d2h: bool, steps: int):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        frames = samples = 0
        for _ in range(steps):
            f, s = run_pass(no_d2h)
            frames += f
            samples += s
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        return sh.reduce_timing(dist if world > 1 else None, torch, "cuda", dev_ms, float(frames), float(samples),
                                wall * 1e3)

    for _ in range(args.warmup):
        run_pass(True)
    # (1) value: K timed steps, PCM left on the device, no instrumentation
    rt.stage_times(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, frames, _, wall_ms = timed(True, args.steps)
    st = rt.stage_times()
    # (2) e2e: K timed steps through the public call with host buffers
    rt.stage_times(reset=True)
    ms2, frames2, samples2, _ = timed(False, args.steps)
    st_e2e = rt.stage_times()
    clocks = sampler.stop() if rank == 0 else {}
    # (3) one more timed step with per-launch CUDA events (roofline accounting;
    #     event pairs around every GEMM/attention launch, CUDA graphs off)
    eng0 = rt.engine(0) if args.pool else rt  # per-kernel profile and the latency leg run on one engine
    rt.stage_times(reset=True)
    for i in range(n_dev):
        (rt.engine(i) if args.pool else rt).profile(True)
    ms3, frames3, _, _ = timed(True, 1)
    prof = eng0.profile_report()
    st3 = rt.stage_times()
    for i in range(n_dev):
        (rt.engine(i) if args.pool else rt).profile(False)
    value = sh.job_throughput(frames, ms)
    e2e = sh.job_throughput(frames2, ms2)
    out = None
    if rank == 0:
        peaks = load_peaks()
        info = rt.info
        shapes = {k[6:]: {"launches": v["launches"], "ms": round(v["ms"], 3),
                          "tflops": round(v["flops"] / (v["ms"] * 1e9), 2) if v["ms"] else 0.0}
                  for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]) if k.startswith("shape/")}
        prof = {k: v for k, v in prof.items() if not k.startswith("shape/")}
        classes = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
        tot_ms = sum(v["ms"] for _, v in classes) or 1.0
        kc = {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "share_of_profiled": round(v["ms"] / tot_ms, 4),
                  "tflops": round(v["flops"] / (v["ms"] * 1e9), 3) if v["ms"] else 0.0,
                  "gbs": round(v["bytes"] / (v["ms"] * 1e6), 1) if v["ms"] else 0.0} for k, v in classes}
        ridge = peaks["tf_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)  # FLOP per byte
        ncu_traffic = load_ncu_traffic()

        def roofline_of(name, tv):
            """Roofline entry of one kernel class: algorithmic bytes (or flops) / event-timed duration."""
            intensity = tv["flops"] / max(tv["bytes"], 1.0)
            common = {"kernel": name, "per_launch_us": 1e3 * tv["ms"] / max(tv["launches"], 1),
                      "launches": tv["launches"], "share_of_profiled_gpu_time": tv["ms"] / tot_ms,
                      "arithmetic_intensity_flop_per_byte": intensity,
                      "traffic": ncu_traffic.get(name)}
            if intensity < ridge:
                ach = tv["bytes"] / (tv["ms"] * 1e6)
                return dict(common, bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s",
                            frac=ach / peaks["hbm_gbs"], peak_source=peaks["source"],
                            also_tflops=tv["flops"] / (tv["ms"] * 1e9))
            ach = tv["flops"] / (tv["ms"] * 1e9)
            return dict(common, bound="tensor", achieved=ach, peak=peaks["tf_sustained"], unit="TFLOP/s",
                        frac=ach / peaks["tf_sustained"], issued_tflops=3.0 * ach,
                        frac_issued=3.0 * ach / peaks["tf_sustained"],
                        peak_source=peaks["source"] + " (sustained bf16; kernel timed inside a long step)",
                        note="split-bf16 operands: 3 tensor-core products per useful fp32-accurate product "
                             "(issued = 3 x achieved), so the useful ceiling is peak/3")

        # the dominant kernel is chosen by kernel FUNCTION: every tcgen05_* class (tile width / epilogue-warp
        # variants) is the one templated tc_gemm_kernel; the per-tile split stays in roofline_by_class
        def function_of(cls):
            return "tc_gemm_kernel" if cls.startswith("tcgen05_") else cls
        by_fn = {}
        for k, v in classes:
            acc = by_fn.setdefault(function_of(k), {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0, "classes": []})
            for f in ("launches", "ms", "flops", "bytes"):
                acc[f] += v[f]
            acc["classes"].append(k)
        top, tv = max(by_fn.items(), key=lambda kv: kv[1]["ms"])
        roofline = dict(roofline_of(top, tv), classes_merged=tv["classes"],
                        algorithmic="2*M*N*K flops per GEMM launch, 4*(N*K + M*min(K,lda) + M*N) bytes; attention: "
                                    "4*keys*D flops, K+V rows read once (DESIGN.md section 4)",
                        measured_over="one timed step with a CUDA-event pair around every launch on the engine "
                                      "stream (CUDA graphs off); algorithmic flops/bytes per launch summed",
                        profiled_step_ms=ms3)
        roofline_by_class = {k: roofline_of(k, v) for k, v in classes if v["ms"] > 0 and (v["flops"] or v["bytes"])}
        alg_flops = 2.0 * (info.macs_ar_frame + info.macs_mimi_frame) * frames
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus_job, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config,
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d * world),
                    "d2h_bytes_per_step": int(samples2 * 4 / max(args.steps, 1))},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": roofline, "roofline_by_class": roofline_by_class,
            "kernel_classes": kc, "gemm_shapes": shapes,
            "stage_ms_rank0": {k: round(st[k], 2) for k in ("prefill_ms", "step_ms", "flow_head_ms", "mimi_ms", "d2h_ms")},
            "stage_ms_e2e_rank0": {k: round(st_e2e[k], 2)
                                   for k in ("prefill_ms", "step_ms", "flow_head_ms", "mimi_ms", "d2h_ms")},
            "stage_note": "device time per stage over the K timed steps (CUDA events; flow_head_ms is the part of "
                          "step_ms after the transformer, d2h_ms the PCM copies on the copy stream, which overlap "
                          "compute); the value leg leaves PCM on the device, so its d2h_ms is 0 by construction; "
                          "mimi_ms runs on its own stream beside step_ms, so the stages do not add up to the wall "
                          "time, and with programmatic dependent launch the in-graph event that splits off "
                          "flow_head_ms sits at a full (non-programmatic) dependency, so the split is approximate",
            "mean_batch_occupancy": st["step_rows"] / max(st["steps"], 1),
            "steps_per_s_rank0": st["steps"] / (ms / 1e3),
            "realtime_factor_per_gpu": value / n_gpus_job,
            "achieved_tflops_job": alg_flops / (ms * 1e9),
            "frac_of_tensor_roofline_job": alg_flops / (ms * 1e9) / (peaks["tf_sustained"] * n_gpus_job),
            "wall_ms_per_step": wall_ms / max(args.steps, 1),
            "frames_per_step": frames / max(args.steps, 1),
        }

    # ---------------- latency leg: C2 first chunk, B=1, streaming-style -------
    if rank == 0 and not args.no_latency:
        c2 = wl.paragraph_200_words()[0]
        n2 = c2.noise()
        ttfa = []
        for i in range(12):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s = eng0.seq_open(c2.tokens, reserve_frames=8)
            eng0.step([s], n2[0])
            pcm = s.decode(0, 1)
            ttfa.append((time.perf_counter() - t0) * 1e3)
            s.close()
        ttfa = sorted(ttfa[2:])
        s = eng0.seq_open(c2.tokens, reserve_frames=256)
        for i in range(8):
            eng0.step([s], n2[i])
        eng0.stage_times(reset=True)
        t0 = time.perf_counter()
        nb1 = 64
        for i in range(nb1):
            eng0.step([s], n2[8 + i])
        b1_wall = (time.perf_counter() - t0) / nb1
        st1 = eng0.stage_times()
        eng0.profile(True)   # separate instrumented steps for the per-kernel GB/s (graphs off)
        for i in range(16):
            eng0.step([s], n2[80 + i])
        p1 = eng0.profile_report()
        eng0.profile(False)
        s.close()
        # C2 as written: the 200-word paragraph (6 chunks) streamed at batch 1 through the public
        # scheduler call with PTTS_REQ_STREAM (one delivery per frame): every frame's PCM is on the host
        # when ptts_ticket_read announces it, so the intervals below include AR step + Mimi decode + D2H.
        c2_first, c2_gaps, c2_audio, c2_wall = [], [], 0.0, 0.0
        para = wl.paragraph_200_words()
        for rep in range(3):
            for c in para:
                cfg = pkg.RuntimeGenerateConfig(max_steps=c.max_steps, eos_threshold=1e30, force_eos_step=c.force_eos_step,
                                                frames_after_eos=c.frames_after_eos, noise=c.noise(), stream_frames=1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                pend = eng0.submit(c.tokens, cfg)
                have, last = 0, t0
                while True:
                    ready, done = eng0.read(pend, have)
                    now = time.perf_counter()
                    if ready > have:
                        if have == 0:
                            c2_first.append((now - t0) * 1e3)
                        else:
                            c2_gaps.extend([(now - last) * 1e3 / (ready - have)] * (ready - have))
                        have, last = ready, now
                    if done:
                        break
                r = eng0.wait(pend)
                if rep > 0:  # first repetition warms the graphs of the B=1 path
                    c2_audio += r.n_frames * 0.08
                    c2_wall += time.perf_counter() - t0
        c2_first, c2_gaps = sorted(c2_first[len(para):]), sorted(c2_gaps)
        c2s = {"workload": "C2: 200-word paragraph, %d chunks of <= 50 tokens, batch 1, streamed per frame" % len(para),
              "first_audio_ms_p50": c2_first[len(c2_first) // 2], "first_audio_ms_max": c2_first[-1],
              "frame_emit_ms_p50": c2_gaps[len(c2_gaps) // 2], "frame_emit_ms_p99": c2_gaps[int(len(c2_gaps) * 0.99)],
              "realtime_factor_including_decode": c2_audio / c2_wall,
              "definition": "host clock; first_audio = submit -> PCM of frame 0 on the host (prefill + 1 AR frame + "
                            "Mimi decode + D2H); frame_emit = interval between consecutive per-frame deliveries; every "
                            "delivery decodes the new frame + one context frame against the request's Mimi K/V ring "
                            "(PTTS_NO_STREAM_STATE=1: the 33-frame receptive field instead)"}
        gv = p1.get("gemv_f32", {"ms": 0, "bytes": 0, "launches": 1})
        peaks = load_peaks()
        bytes_frame = 4.0 * info.params_ar
        out["latency"] = {
            "workload": "C2 first chunk (%d tokens), batch 1" % c2.tokens.size,
            "ttfa_ms_p50": ttfa[len(ttfa) // 2], "ttfa_ms_min": ttfa[0],
            "ttfa_definition": "seq_open (prefill) + 1 AR frame + Mimi decode of frame 0 + PCM D2H, host clock",
            "b1_frame_us_device": 1e3 * st1["step_ms"] / nb1, "b1_frame_us_wall": 1e6 * b1_wall,
            "b1_realtime_factor": 0.08 / b1_wall,
            "b1_weight_stream_gbs": bytes_frame / (st1["step_ms"] / nb1 * 1e6),
            "b1_frac_of_hbm_roofline": bytes_frame / (st1["step_ms"] / nb1 * 1e6) / peaks["hbm_gbs"],
            "gemv_kernel_gbs": gv["bytes"] / (gv["ms"] * 1e6) if gv["ms"] else None,
            "launches_per_frame": st1["kernel_launches"] / nb1, "cuda_graph": True,
            "c2_stream": c2s,
        }

    # ---------------- CPU baseline beside it (rank 0, N=1 only) ---------------
    if rank == 0 and world == 1 and not args.pool and not args.no_cpu_baseline:
        o = ge.oracle()
        cores = os.cpu_count() or 1
        om = o.OracleModel(tensors)
        sample_chunks = wl.make_workload(cores, 5, 60, seed=7)[:cores]
        a, t, nfr = cpu_run(om, sample_chunks, args.cpu_sample_frames, cores)
        cap_txt = f"capped at {args.cpu_sample_frames} frames each" if args.cpu_sample_frames > 0 else "whole chunks"
        out["cpu_baseline"] = {
            "value": a / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(sample_chunks)} chunks (first of the workload), {cap_txt} ({nfr} frames), one chunk per "
                      f"host thread, {t:.1f} s of wall time",
            "note": "C restatement of the reference path (oracle/); the Go reference cannot be built here"}
    if rank == 0:
        emit(json.dumps(out))
    rt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
