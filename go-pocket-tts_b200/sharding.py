"""Multi-GPU sharding of the synthesis workload.

Utterance chunks are independent (reference internal/tts/service.go:138-153
generates each chunk with fresh state), so the path shards by utterance with NO
data-path collective: one process per GPU, a full replica of the weights, its
own KV pool and scheduler.  The only cross-rank traffic is the benchmark's
timing reduction (max of elapsed time, sum of audio produced)."""
from __future__ import annotations


def shard_range(rank: int, world: int, per_gpu: int):
    """Weak-scaling shard: rank r owns utterances [r*per_gpu, (r+1)*per_gpu)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return rank * per_gpu, (rank + 1) * per_gpu


def reduce_timing(dist, torch, device, elapsed_ms: float, frames: float, samples: float, wall_ms: float):
    """MAX over ranks for times, SUM for work.  Returns (ms, frames, samples, wall_ms)."""
    t = torch.tensor([elapsed_ms, frames, samples, wall_ms], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        return tmax[0].item(), tsum[1].item(), tsum[2].item(), tmax[3].item()
    return t[0].item(), t[1].item(), t[2].item(), t[3].item()


def job_throughput(frames_total: float, elapsed_ms_max: float) -> float:
    """audio-seconds per wall-second of the whole job."""
    return (frames_total * 1920.0 / 24000.0) / (elapsed_ms_max / 1e3)
