"""Render-job driver: one process per GPU, samples sharded across ranks, one sum-reduce of the accumulators.

Partition: rank r of N renders ALL pixels for the global sample indices
``[r*spp/N, (r+1)*spp/N)``.  Philox counters carry the global sample index, so the union over ranks is exactly
the sample set a single GPU would draw; the only exchange is a float32 sum of the W*H*3 accumulators
(``torch.distributed.reduce`` — NCCL over NVLink on GPUs, gloo in the CPU tests of this host logic).

PyTorch is used here for device memory, streams and the process group only; every pixel is computed by
``libare_b200.so`` through :mod:`aurora_rendering_engine_b200.capi`.
"""
from __future__ import annotations

import hashlib
import os
from dataclasses import dataclass

import numpy as np

from . import capi


def shard_samples(spp: int, world: int, rank: int) -> tuple[int, int]:
    """(first global sample, count) of ``rank``; remainders go to the low ranks; disjoint and covering."""
    if world < 1 or not (0 <= rank < world) or spp < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(spp, world)
    count = base + (1 if rank < rem else 0)
    begin = rank * base + min(rank, rem)
    return begin, count


def chunk_ranges(begin: int, count: int, chunk: int):
    """Split [begin, begin+count) into launches of at most ``chunk`` samples per pixel."""
    if chunk < 1:
        raise ValueError("chunk must be >= 1")
    out = []
    s = begin
    while s < begin + count:
        c = min(chunk, begin + count - s)
        out.append((s, c))
        s += c
    return out


def merge_ranges(ranges):
    """Union of half-open sample ranges given as (begin, count), sorted and coalesced."""
    iv = sorted((int(b), int(b) + int(c)) for b, c in ranges if c > 0)
    out = []
    for b, e in iv:
        if out and b <= out[-1][1]:
            out[-1][1] = max(out[-1][1], e)
        else:
            out.append([b, e])
    return [(b, e - b) for b, e in out]


def missing_ranges(begin: int, count: int, done):
    """Sample ranges of [begin, begin+count) NOT covered by ``done`` — what a resumed job, or the rank that takes over
    from a failed one, still has to render.  The Philox counter is the GLOBAL sample index, so any rank can render any
    range and the result is the same set of samples."""
    out, s, end = [], begin, begin + count
    for b, c in merge_ranges(done):
        e = b + c
        if e <= s:
            continue
        if b >= end:
            break
        if b > s:
            out.append((s, b - s))
        s = max(s, e)
    if s < end:
        out.append((s, end - s))
    return out


def overlapping(ranges) -> bool:
    """True when two ranges share a sample (a double-counted sample would bias the image)."""
    iv = sorted((int(b), int(b) + int(c)) for b, c in ranges if c > 0)
    return any(iv[i][0] < iv[i - 1][1] for i in range(1, len(iv)))


def dist_env():
    """(rank, local_rank, world) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def reduce_sum_to_root(tensor, world: int, root: int = 0):
    """In-place sum-reduce of the accumulator onto ``root`` (no-op for one rank)."""
    if world > 1:
        import torch.distributed as dist
        dist.reduce(tensor, dst=root, op=dist.ReduceOp.SUM)
    return tensor


@dataclass
class JobResult:
    accum: object          # torch tensor (H,W,3) float32 of sample sums (valid on root after finish())
    samples: int = 0       # this rank
    rays: int = 0          # this rank (only counted when stats are requested)
    kernel_ms: float = 0.0
    launches: int = 0


class RenderJob:
    """A scene committed on this rank's GPU plus a device accumulator; ``render_range`` adds samples into it."""

    def __init__(self, scene_desc, device_index: int = 0, traversal: int = 0, options: dict | None = None, builder: int | None = None):
        import torch
        self.torch = torch
        self.sc = scene_desc
        self.device_index = device_index
        torch.cuda.set_device(device_index)
        self.ctx = capi.Context(device_index)
        self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        for opt, val in (options or {}).items():   # capi.OPT_* -> value (are_cuda_set_option)
            self.ctx.set_option(opt, val)
        if builder is not None:
            self.ctx.set_bvh_builder(builder)
        scene_desc.feed(self.ctx)
        self.h2d_bytes = self.ctx.commit()
        self.cam = capi.make_camera(**scene_desc.camera_args())
        self.traversal = traversal
        self.accum = torch.zeros((scene_desc.height, scene_desc.width, 3), dtype=torch.float32, device=f"cuda:{device_index}")
        self.result = JobResult(self.accum)
        self.done = []  # (begin, count) sample ranges already in the accumulator
        self.shard = None  # (begin, count) this job is responsible for (render_job sets it); None = unrestricted
        self.rank, self.world = 0, 1

    def params(self, begin, count, **over):
        return capi.make_params(**self.sc.params_args(sample_begin=begin, sample_count=count, traversal=self.traversal, **over))

    def render_range(self, begin: int, count: int, want_stats: bool = False, **over):
        st = self.ctx.render_device(self.cam, self.params(begin, count, **over), self.accum.data_ptr(), want_stats=want_stats)
        self.result.launches += 1
        self.result.samples += self.sc.width * self.sc.height * count
        self.done.append((int(begin), int(count)))
        if st is not None:
            self.result.rays += st.rays
            self.result.kernel_ms += st.kernel_ms
        return st

    def finish(self, world: int):
        reduce_sum_to_root(self.accum, world)
        return self.result

    def image(self, total_spp: int) -> np.ndarray:
        return (self.accum / float(total_spp)).cpu().numpy()

    def save_ppm(self, path: str, total_spp: int, encoder: int = 0):
        rgb8 = self.ctx.tonemap(self.accum.data_ptr(), self.sc.width, self.sc.height, 1.0 / total_spp, encoder)
        self.ctx.write_ppm(path, rgb8)
        return rgb8

    # -- checkpoint / resume: the accumulator holds plain sample sums and the RNG is counter-based, so a job is fully
    #    described by (accumulator, which global sample ranges are in it)
    def signature(self) -> str:
        """Everything that decides WHICH estimator the accumulator holds: scene, frame, depth, integrator, traversal, and a
        digest of the camera and of every render parameter except the sample range (seed, t_min, background, ...).  A
        checkpoint written under another signature holds samples of another picture."""
        sc = self.sc
        par = {k: v for k, v in sc.params_args(traversal=self.traversal).items() if k not in ("sample_begin", "sample_count")}
        blob = repr((sorted((k, repr(v)) for k, v in sc.camera_args().items()), sorted((k, repr(v)) for k, v in par.items())))
        return (f"{sc.name}|{sc.width}x{sc.height}|{sc.num_prims}|depth{sc.max_depth}|integrator{sc.integrator}|"
                f"{hashlib.sha256(blob.encode()).hexdigest()[:16]}")

    def save_checkpoint(self, path: str):
        self.torch.cuda.current_stream().synchronize()
        tmp = path + ".tmp.npz"
        shard = self.shard if self.shard is not None else (-1, -1)
        np.savez(tmp, accum=self.accum.cpu().numpy(), done=np.array(merge_ranges(self.done), dtype=np.int64).reshape(-1, 2),
                 signature=np.array(self.signature()), shard=np.array(shard, dtype=np.int64), rank_world=np.array([self.rank, self.world], dtype=np.int64))
        os.replace(tmp, path)

    def load_checkpoint(self, path: str):
        """Restore accumulator + progress.  Raises ValueError for a checkpoint of another job (scene, camera, seed, ...), of
        another rank / world size / shard, or one whose recorded sample ranges overlap or leave this job's shard — loading
        any of those would put samples into the sum twice or samples of another estimator beside this one's."""
        ck = np.load(path)
        if str(ck["signature"]) != self.signature():
            raise ValueError(f"checkpoint belongs to {ck['signature']}, this job is {self.signature()}")
        done = [(int(b), int(c)) for b, c in ck["done"]]
        if overlapping(done):
            raise ValueError("checkpoint lists overlapping sample ranges")
        if self.shard is not None:
            if "shard" in ck.files and tuple(int(x) for x in ck["shard"]) not in ((-1, -1), tuple(self.shard)):
                raise ValueError(f"checkpoint was written for shard {tuple(int(x) for x in ck['shard'])} "
                                 f"(rank/world {tuple(int(x) for x in ck['rank_world'])}), this rank renders {tuple(self.shard)}")
            lo, hi = self.shard[0], self.shard[0] + self.shard[1]
            for b, c in done:
                if b < lo or b + c > hi:
                    raise ValueError(f"checkpoint holds samples [{b}, {b + c}) outside this rank's shard [{lo}, {hi})")
        self.accum.copy_(self.torch.from_numpy(ck["accum"]))
        self.done = done
        return self.done

    def close(self):
        self.ctx.close()


def render_job(scene_desc, spp: int, chunk: int = 64, traversal: int = 0, want_stats: bool = False, checkpoint: str | None = None,
               checkpoint_every: int = 0):
    """Whole job on this rank (call under torchrun for N GPUs): shard, render in chunks, reduce. Returns RenderJob.
    With ``checkpoint`` (a path; ``{rank}`` is substituted) the rank resumes from that file when it exists and rewrites
    it every ``checkpoint_every`` launches (0 = only at the end of its shard)."""
    import torch
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if checkpoint and world > 1 and "{rank}" not in checkpoint:
        raise ValueError("a multi-rank job needs one checkpoint file per rank: put {rank} in the path")
    job = RenderJob(scene_desc, local_rank, traversal)
    begin, count = shard_samples(spp, world, rank)
    job.shard, job.rank, job.world = (begin, count), rank, world
    ck = checkpoint.format(rank=rank) if checkpoint else None
    if ck and os.path.exists(ck):
        job.load_checkpoint(ck)
    n = 0
    for mb, mc in missing_ranges(begin, count, job.done):
        for b, c in chunk_ranges(mb, mc, chunk):
            job.render_range(b, c, want_stats=want_stats)
            n += 1
            if ck and checkpoint_every and n % checkpoint_every == 0:
                job.save_checkpoint(ck)
    if ck:
        job.save_checkpoint(ck)
    job.finish(world)
    return job
