#!/usr/bin/env python
"""bench.py — throughput of the path-tracing hot path on the BASELINE.json headline workload.

Workload (config.workload): BASELINE config 3 — Cornell box (5 walls + light + 2 boxes as 36 triangles, 555-unit
dimensions), 2048x2048, max depth 50.  The job's 16384 spp are rendered in chunks; ONE STEP = one chunk of
``--spp-per-step`` samples per pixel over the whole frame on every rank (weak scaling: rank r, step k renders
global samples [(k*N + r)*S, +S)), and the timed region ends with the job's NCCL sum-reduce of the accumulators.

  value      Msamples/s, whole job, scene + accumulator resident in HBM, CUDA events on the launch stream,
             max over ranks.  (Mrays/s is reported beside it.)
  e2e        the same metric through the public host-buffer API.  N = 1: are_cuda_commit (scene H2D) + are_cuda_render
             (zero, render, W*H*3 float D2H into a PAGEABLE host buffer) every step.  N > 1: the real job step — every
             rank commits and renders its chunk, one NCCL sum-reduce, rank 0 copies the summed frame to pageable host memory.
  roofline   the render kernel against the FP32 FMA issue peak (this path has no dense contraction and its
             working set lives in shared memory, so neither HBM nor tensor peak bounds it — DESIGN.md §5);
             achieved = counted primitive tests x SURVEY §8d FLOP figures / kernel time; traffic = dram bytes of
             one launch measured by an ncu side-run of this very command (N = 1; null when ncu is unavailable).
  cpu_baseline  the CPU twin (oracle/are_oracle.c built -O3 -march=native on this box, "port": the reference has no
             renderer for this config) on all host threads, on a bounded sample of the same frame.
  strong_job the job the north_star names — Cornell box at ``--job-spp`` (1024) samples per pixel, a FIXED amount of work
             sharded over the N GPUs — timed from the first launch to the tone-mapped P6 bytes on rank 0 (render, reduce,
             k_tonemap, D2H inside): strong scaling beside the weak-scaling headline.
  reduce_check (N > 1) the N-rank sum against the same global samples rendered by rank 0 alone (>= 60 dB).
  configs    the other BASELINE configurations (0, 1, 2, 4) measured the same way in short steps, each with its own
             roofline fraction and (N = 1) CPU baseline; config 0's CPU column is the REAL reference program
             (oracle/_ref/rt_ref = experiments/rt.cpp unmodified, one thread).

``--impl reference`` times the CPU implementation alone (rank 0 only under torchrun): the fast build of the twin for the
headline config, the real reference program for ``--scene rt_cornell``.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Msamples/s (Cornell box 2048x2048, depth 50; Mrays/s alongside)"
UNIT = "Msamples/s"

# SURVEY.md §8d algorithmic FLOP figures
F_TRI, F_SPH, F_QUAD, F_AABB, F_SCATTER, F_PRIMARY = 51.0, 28.0, 45.0, 24.0, 40.0, 40.0
# A parallelepiped test stands for the up-to-six parallelogram tests it replaces; SURVEY §8d has no figure for it.  The
# arithmetic it requires: 3 x (two 3-term dots 10 + rcp 1 + 2 mul + 2 add) + 4 min/max = 49 FLOP.
F_BOX = 49.0
# SURVEY §8d algorithmic bytes of a traversal: 64-byte BVH2 node per visit (the survey's 32 B per AABB x 2 children),
# 48-byte plane-form record per triangle / quad test, 32 bytes per sphere test (centre, radius, r^2), 96 per box
B_NODE, B_PLANE, B_SPHERE, B_BOXREC = 64.0, 48.0, 32.0, 96.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="cornell_box")
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--height", type=int, default=2048)
    ap.add_argument("--spp-per-step", type=int, default=256)
    ap.add_argument("--traversal", type=int, default=0)
    ap.add_argument("--kernel", default="auto", choices=["auto", "lean", "generic", "baked-packed"],
                    help="A/B: auto = scene-specialised (baked) kernel where the scene has a lean form; lean = precompiled lean kernel; generic = generic brute-force kernel")
    ap.add_argument("--builder", type=int, default=-1, help="BVH builder: 0 host SAH, 1 device LBVH (default: host; device for the 1 M-primitive scene)")
    ap.add_argument("--baked-min-blocks", type=int, default=0, help="tuning: CTAs per SM the baked kernel is compiled for")
    ap.add_argument("--wavefront", action="store_true", help="A/B: the same estimator scheduled as wavefront stages (ARE_INTEGRATOR_PATH_WAVEFRONT)")
    ap.add_argument("--l2-persist", type=int, default=0, help="A/B: BVH renders mark the node array L2-persisting, per cent of the carve-out (ARE_OPT_L2_PERSIST_NODES)")
    ap.add_argument("--n-prims", type=int, default=0, help="A/B: primitive count of the `stress` scene (default: its 1 000 000)")
    ap.add_argument("--no-quant", action="store_true", help="A/B: big hierarchies through their fp32 nodes (ARE_OPT_QUANTIZED_NODES = 0)")
    ap.add_argument("--job-spp", type=int, default=1024, help="strong-scaling job: total samples per pixel sharded over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of BASELINE configs 0, 1, 2, 4")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu side-run that measures dram bytes per launch")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--traffic-probe", action="store_true", help=argparse.SUPPRESS)  # internal: the command ncu profiles
    return ap.parse_args()


def make_scene(a, name=None, **kw):
    from aurora_rendering_engine_b200 import scenes
    if name is None:
        kw = {"n_prims": a.n_prims} if a.scene == "stress" and a.n_prims > 0 else {}
        sc = scenes.by_name(a.scene, width=a.width, height=a.height, **kw)
        if a.wavefront:
            sc.integrator = scenes.INTEGRATOR_PATH_WAVEFRONT
        return sc
    return scenes.by_name(name, **kw)


BASELINE_CONFIG = {"rt_cornell": "0; reference program: 1 primary + 32 AO rays per pixel", "rtiow_final": "1; full job = 500 spp",
                   "textured": "2; full job = 1024 spp", "cornell_box": "3; full job = 16384 spp", "stress": "4; full job = 256 spp"}
# (scene, spp per step) of the short runs in `configs`
OTHER_CONFIGS = (("rt_cornell", 1), ("rtiow_final", 100), ("textured", 256), ("stress", 16))


def workload_name(sc, S):
    return (f"{sc.name} {sc.width}x{sc.height} depth {sc.max_depth} ({sc.num_prims} primitives), "
            f"{S} spp per step per GPU (BASELINE config {BASELINE_CONFIG.get(sc.name, '?')})")


def workload_config(a, sc, world):
    return {
        "workload": workload_name(sc, a.spp_per_step),
        "spp_per_step": a.spp_per_step, "width": sc.width, "height": sc.height, "max_depth": sc.max_depth,
        "parallelism": f"sample-sharded x{world}, one NCCL sum-reduce of the accumulators at job end",
        "l2": "flushed between timed steps (256 MiB device write); per-step CUDA events summed",
    }


# ---------------------------------------------------------------------------------------------------------
# CPU implementations — used for cpu_baseline and for --impl reference
# ---------------------------------------------------------------------------------------------------------
def fast_oracle_path():
    """oracle/liboracle_fast.so: the twin rebuilt ON THIS BOX with -O3 -march=native."""
    r = subprocess.run(["make", "-B", "-C", os.path.join(ROOT, "oracle"), "fast"], capture_output=True, text=True)
    p = os.path.join(ROOT, "oracle", "liboracle_fast.so")
    return p if r.returncode == 0 and os.path.exists(p) else None


class CpuTwin:
    def __init__(self, sc):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_binding import Oracle
        from aurora_rendering_engine_b200 import capi
        self.capi = capi
        self.sc = sc
        fast = fast_oracle_path()
        self.build = "-O3 -march=native (oracle/liboracle_fast.so, built on this box)" if fast else "-O2 (oracle/liboracle.so; the fast build failed)"
        self.orc = Oracle(fast)
        self.osc = sc.feed(self.orc.scene())
        self.cam = capi.make_camera(**sc.camera_args())
        self.threads = os.cpu_count() or 1
        if sc.num_prims > 256:  # the twin builds its AABB tree on the first query: keep that out of every timing
            self.run(1, rows=(0, 1))

    def run(self, spp, rows=None, sample_begin=0):
        """Render `spp` samples of rows [r0,r1) (default: whole frame). Returns (seconds, samples, rays)."""
        W, H = self.sc.width, self.sc.height
        r0, r1 = rows if rows else (0, H)
        par = self.capi.make_params(**self.sc.params_args(sample_begin=sample_begin, sample_count=spp))
        t0 = time.perf_counter()
        _, st = self.osc.render(self.cam, par, nthreads=self.threads, window=(0, r0, W, r1))
        dt = time.perf_counter() - t0
        return dt, int(st.samples), int(st.rays)

    def calibrate(self, target_s):
        """Pick (spp, rows) so that one run takes about target_s: a probe of a row band first."""
        W, H = self.sc.width, self.sc.height
        band = max(1, H // 64)
        dt, n, _ = self.run(1, rows=(H // 2 - band // 2, H // 2 - band // 2 + band))
        rate = n / max(dt, 1e-6)  # samples/s
        frame = W * H
        spp = int(rate * target_s / frame)
        if spp >= 1:
            return min(spp, 64), None, rate
        rows = max(band, int(rate * target_s / W))
        r0 = max(0, H // 2 - rows // 2)
        return 1, (r0, min(H, r0 + rows)), rate


def cpu_baseline(sc, seconds):
    if sc.name == "rt_cornell":
        return rt_reference_baseline()
    twin = CpuTwin(sc)
    spp, rows, _ = twin.calibrate(seconds)
    dt, n, rays = twin.run(spp, rows)
    what = f"{spp} spp of " + ("the whole frame" if rows is None else f"rows {rows[0]}..{rows[1]} of the frame")
    return {"value": n / dt / 1e6, "unit": UNIT, "cores": twin.threads, "kind": "port",
            "sample": f"{what} ({n} samples, {dt:.1f} s), fp64 CPU twin oracle/are_oracle.c {twin.build}, pthreads over rows",
            "mrays_per_s": rays / dt / 1e6}


def rt_reference_baseline():
    """Config 0: the REAL reference program — oracle/_ref/rt_ref = experiments/rt.cpp compiled unmodified (-O2, its own
    flags), single-threaded, its one hard-wired job (512x512, 1 primary + 32 AO rays per pixel); true ray count from the
    sed-instrumented twin build rt_ref_counted."""
    ref = os.path.join(ROOT, "oracle", "_ref", "rt_ref")
    cnt = os.path.join(ROOT, "oracle", "_ref", "rt_ref_counted")
    if not os.path.exists(ref):
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref/rt_ref not present on this box"}
    d = tempfile.mkdtemp()
    t0 = time.perf_counter()
    subprocess.run([ref], cwd=d, capture_output=True, check=True, timeout=600)
    dt = time.perf_counter() - t0
    rays = None
    if os.path.exists(cnt):
        r = subprocess.run([cnt], cwd=d, capture_output=True, timeout=600, env=dict(os.environ, ARE_RT_SEED="1"))
        m = re.search(rb"ARE_COUNT rays=(\d+)", r.stdout)
        rays = int(m.group(1)) if m else None
    return {"value": 512 * 512 / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"the whole job of experiments/rt.cpp (oracle/_ref/rt_ref, unmodified, g++ -O2, 1 thread): 512x512, 1 sample per pixel, {dt:.2f} s",
            "mrays_per_s": rays / dt / 1e6 if rays else None, "rays": rays}


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sc = make_scene(a)
    cfg = workload_config(a, sc, 1)
    if sc.name == "rt_cornell":
        runs = [rt_reference_baseline() for _ in range(max(1, min(a.steps, 3)))]
        val = sum(r["value"] for r in runs) / len(runs)
        base = dict(runs[-1], value=val)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": len(runs), "warmup": 0,
                "ms_per_step": 512 * 512 / val / 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg, "mrays_per_s": base.get("mrays_per_s"), "cpu_baseline": base,
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return
    twin = CpuTwin(sc)
    budget = 150.0 / max(1, a.steps + a.warmup)
    spp, rows, _ = twin.calibrate(min(10.0, budget))
    for k in range(a.warmup):
        twin.run(spp, rows, sample_begin=k * spp)
    tot_t = tot_n = tot_r = 0
    for k in range(a.steps):
        dt, n, r = twin.run(spp, rows, sample_begin=(a.warmup + k) * spp)
        tot_t += dt; tot_n += n; tot_r += r
    val = tot_n / tot_t / 1e6
    what = f"{spp} spp of " + ("the whole frame" if rows is None else f"rows {rows[0]}..{rows[1]}") + " per step"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": tot_t / max(1, a.steps) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "mrays_per_s": tot_r / tot_t / 1e6,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": twin.threads, "kind": "port",
                             "sample": what + f"; the reference ships no renderer for this config (SURVEY.md §0), so its CPU path is the fp64 twin in oracle/are_oracle.c, {twin.build}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            hi = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
            out = {"sm_mhz": sorted(hi)[len(hi) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def kernel_name(capi, variant, baked_kind=1):
    if variant == capi.KERNEL_BRUTE_BAKED and baked_kind == 2:
        return "k_render_baked (generic brute-force kernel, scene compiled in by NVRTC)"
    return {capi.KERNEL_RT_AO: "k_render_rtao", capi.KERNEL_BRUTE: "k_render_path<brute/smem>",
            capi.KERNEL_BRUTE_LEAN: "k_render_path<brute/smem, lean>",
            capi.KERNEL_BRUTE_BAKED: "k_render_baked (lean kernel, scene compiled in by NVRTC)",
            capi.KERNEL_BVH2: "k_render_path<bvh2>", capi.KERNEL_BVH2_BIG: "k_render_path<bvh2, 12 CTAs/SM>", capi.KERNEL_BVH2_QUANT: "k_render_path<bvh2, quantised 32-byte nodes, 12 CTAs/SM>",
            capi.KERNEL_WIDE: "k_render_path<wide bvh>", capi.KERNEL_BVH4: "k_render_path<bvh4>", capi.KERNEL_WAVEFRONT: "k_wf_generate + k_wf_extend<bvh2> + k_wf_shade (wavefront)"}.get(variant, "?")


class Bench:
    """Everything one scene needs on this rank: a committed RenderJob + timing helpers."""

    def __init__(self, a, sc, S, options=None, builder=None):
        import torch
        import torch.distributed as dist
        from aurora_rendering_engine_b200 import capi, engine
        self.torch, self.dist, self.capi, self.engine = torch, dist, capi, engine
        self.rank, self.local_rank, self.world = engine.dist_env()
        self.sc, self.S = sc, S
        t0 = time.perf_counter()
        self.job = engine.RenderJob(sc, self.local_rank, a.traversal, options=options, builder=builder)
        self.commit_s = time.perf_counter() - t0
        self.ctx = self.job.ctx

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def sample_base(self, step):  # global sample index of this rank's chunk in `step`
        return (step * self.world + self.rank) * self.S

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(self, steps, warmup, flush):
        """W warm-up chunks, then `steps` timed chunks (per-step CUDA events on the launch stream, L2 flushed between
        them) + the job's sum-reduce.  Returns dict(value, total_ms, step_ms, reduce_ms)."""
        torch, job = self.torch, self.job
        for k in range(warmup):
            job.render_range(self.sample_base(k), self.S, want_stats=True)
        if self.world > 1:  # warm the communicator
            self.dist.all_reduce(torch.zeros(1, device="cuda"))
        job.accum.zero_()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        red0, red1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        for k in range(steps):
            if flush is not None:
                flush.fill_(k & 0xFF)  # L2 flush, outside the per-step event pair
            ev[k][0].record()
            job.ctx.render_device(job.cam, job.params(self.sample_base(warmup + k), self.S), job.accum.data_ptr())
            ev[k][1].record()
        red0.record()
        self.engine.reduce_sum_to_root(job.accum, self.world)
        red1.record()
        self.barrier()
        step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
        reduce_ms = red0.elapsed_time(red1)
        total_ms = self.max_over_ranks(sum(step_ms) + reduce_ms)
        samples_all = float(self.sc.width) * self.sc.height * self.S * steps * self.world
        return {"value": samples_all / (total_ms * 1e-3) / 1e6, "total_ms": total_ms, "step_ms": step_ms, "reduce_ms": reduce_ms}

    def counted_chunk(self, step):
        """Exact per-chunk counters (untimed re-run with the counting kernel) for Mrays/s + roofline."""
        scratch = self.torch.zeros_like(self.job.accum)
        return self.ctx.render_device(self.job.cam, self.job.params(self.sample_base(step), self.S), scratch.data_ptr(), want_stats=True, count_tests=True)

    def e2e(self, steps, warmup, recommit=True):
        """End to end through host buffers (pageable, like include/are_cuda.hpp's std::vector)."""
        import numpy as np
        job, ctx, sc = self.job, self.ctx, self.sc
        W, H, S = sc.width, sc.height, self.S
        h2d = [job.h2d_bytes]
        if self.world == 1:
            host = np.empty((H, W, 3), np.float32)  # pageable

            def step(k):
                if recommit:
                    h2d[0] = ctx.commit()
                ctx.render(job.cam, job.params(self.sample_base(k), S), out=host)
            api = ("are_cuda_commit + " if recommit else "") + "are_cuda_render (pageable host buffer)"
        else:
            host_t = self.torch.empty((H, W, 3), dtype=self.torch.float32) if self.rank == 0 else None  # pageable, reused like a host program's frame buffer

            def step(k):
                if recommit:
                    h2d[0] = ctx.commit()
                job.accum.zero_()
                ctx.render_device(job.cam, job.params(self.sample_base(k), S), job.accum.data_ptr())
                self.engine.reduce_sum_to_root(job.accum, self.world)
                if self.rank == 0:
                    host_t.copy_(job.accum)  # the summed frame into pageable host memory (synchronises)
            api = ("are_cuda_commit + " if recommit else "") + "are_cuda_render_device per rank + NCCL sum-reduce + one D2H of the summed frame on rank 0 (pageable)"
        for k in range(warmup):
            step(k)
        self.barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            step(warmup + k)
        self.barrier()
        dt_ms = self.max_over_ranks((time.perf_counter() - t0) * 1e3)
        d2h = W * H * 3 * 4
        return {"value": float(W) * H * S * steps * self.world / (dt_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d[0]) if recommit else 0,
                "d2h_bytes_per_step": int(d2h), "steps": steps, "api": api, "timer": "host clock between barrier + device synchronize, max over ranks"}

    def roofline(self, st, kernel_ms, peaks):
        """FP32-issue roofline of one launch from its exact counters; for BVH kernels also the L2 roofline."""
        flops = (F_TRI * st.tri_tests + F_SPH * st.sphere_tests + F_QUAD * st.quad_tests + F_BOX * st.box_tests + F_AABB * 2 * st.node_visits
                 + F_SCATTER * max(0, st.rays - st.samples) + F_PRIMARY * st.samples)
        achieved = flops / (kernel_ms * 1e-3) / 1e12
        roof = {"bound": "fp32", "achieved": achieved, "peak": peaks["fp32_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["fp32_tflops"] if peaks["fp32_tflops"] else None, "traffic": None,
                "kernel": kernel_name(self.capi, st.kernel_variant, self.ctx.commit_info().baked), "kernel_ms": kernel_ms, "flops_per_launch": flops,
                "counted": {"rays": st.rays, "tri_tests": st.tri_tests, "quad_tests": st.quad_tests, "sphere_tests": st.sphere_tests,
                            "box_tests": st.box_tests, "node_visits": st.node_visits}}
        if st.node_visits > 0 and peaks.get("l2_gbs"):
            b_node = 32 if st.kernel_variant == self.capi.KERNEL_BVH2_QUANT else B_NODE  # quantised nodes: one 32-byte record per visit
            byt = b_node * st.node_visits + B_PLANE * (st.tri_tests + st.quad_tests) + B_SPHERE * st.sphere_tests + B_BOXREC * st.box_tests
            gbs = byt / (kernel_ms * 1e-3) / 1e9
            roof["l2"] = {"bound": "l2", "achieved": gbs, "peak": peaks["l2_gbs"], "unit": "GB/s", "frac": gbs / peaks["l2_gbs"],
                          "bytes_per_launch": byt, "node_bytes": b_node,
                          "note": "algorithmic node + primitive record bytes (SURVEY §8d; 32 B per visit through quantised nodes) against the L2 read bandwidth measured by are_cuda_measure_l2_peak"}
        return roof

    def close(self):
        self.job.close()


def strong_job(b, spp_total, chunk=256):
    """The north_star job: Cornell at `spp_total` samples per pixel, sharded over the ranks, to P6 bytes on rank 0."""
    torch, job, ctx, sc = b.torch, b.job, b.ctx, b.sc
    begin, count = b.engine.shard_samples(spp_total, b.world, b.rank)
    chunks = b.engine.chunk_ranges(begin, count, chunk)
    import numpy as np
    header = b"P6\n%d %d\n255\n" % (sc.width, sc.height)
    ppm = bytearray(len(header) + sc.width * sc.height * 3) if b.rank == 0 else None  # the file image: header + payload
    if ppm is not None:
        ppm[:len(header)] = header
        payload = np.frombuffer(ppm, np.uint8, offset=len(header)).reshape(sc.height, sc.width, 3)
    out = {}
    for rep in range(3):  # the first repetition warms up
        job.accum.zero_()
        b.barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t0 = time.perf_counter()
        e[0].record()
        for s0, c in chunks:
            ctx.render_device(job.cam, job.params(s0, c), job.accum.data_ptr())
        e[1].record()
        b.engine.reduce_sum_to_root(job.accum, b.world)
        e[2].record()
        if b.rank == 0:  # k_tonemap + D2H straight into the payload of the P6 image (synchronises)
            ctx.tonemap(job.accum.data_ptr(), sc.width, sc.height, 1.0 / spp_total, 0, out=payload)
        else:
            torch.cuda.synchronize()
        wall_rank = (time.perf_counter() - t0) * 1e3
        b.barrier()
        wall = b.max_over_ranks(wall_rank)
        render_ms, reduce_ms = b.max_over_ranks(e[0].elapsed_time(e[1])), e[1].elapsed_time(e[2])
        if rep == 0:
            continue
        cur = {"wall_ms": wall, "render_ms_max_rank": render_ms, "reduce_ms_rank0": reduce_ms,
               "tonemap_d2h_ms_rank0": max(0.0, wall_rank - e[0].elapsed_time(e[2])) if b.rank == 0 else None, "ppm_bytes": len(ppm) if ppm else None}
        if not out or wall < out["wall_ms"]:
            out = cur
    samples = float(sc.width) * sc.height * spp_total
    out.update({"job": f"{sc.name} {sc.width}x{sc.height}, {spp_total} spp in all, depth {sc.max_depth}, to tone-mapped P6 bytes on rank 0",
                "scaling": "strong", "n_gpus": b.world, "spp_per_gpu": count, "launches_per_gpu": len(chunks), "value": samples / (out["wall_ms"] * 1e-3) / 1e6,
                "unit": UNIT, "timer": "host clock on each rank from the first launch to the P6 bytes (rank 0) / end of its work (others), max over ranks; best of 2"})
    return out


def reduce_check(b, c=16):
    """N > 1: ranks render disjoint chunks of c samples, NCCL sums them on rank 0; rank 0 then renders the same N*c global
    samples alone.  Same samples, different summation order -> PSNR >= 60 dB, here ~1e-6 relative."""
    import numpy as np
    torch, job, ctx = b.torch, b.job, b.ctx
    job.accum.zero_()
    ctx.render_device(job.cam, job.params(b.rank * c, c), job.accum.data_ptr())
    b.engine.reduce_sum_to_root(job.accum, b.world)
    torch.cuda.synchronize()
    out = None
    if b.rank == 0:
        ref = torch.zeros_like(job.accum)
        ctx.render_device(job.cam, job.params(0, c * b.world), ref.data_ptr())
        torch.cuda.synchronize()
        n = float(c * b.world)
        x, y = (job.accum / n).clamp(0, 1), (ref / n).clamp(0, 1)
        mse = float(((x - y).double() ** 2).mean().item())
        diff = float((job.accum - ref).abs().max().item())
        out = {"psnr_db": float(10.0 * np.log10(1.0 / mse)) if mse > 0 else 999.0, "max_abs_diff_of_sums": diff,
               "max_rel": diff / max(1e-30, float(ref.abs().max().item())), "samples_per_pixel": int(n),
               "what": f"{b.world} ranks x {c} spp summed by NCCL vs the same {int(n)} global samples rendered by rank 0 alone, whole frame"}
    b.barrier()
    return out


def traffic_probe(a):
    """ncu side-run of this command (one launch of the headline kernel): dram bytes read + written per launch."""
    ncu = "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:k_render", "-c", "1", "--csv",
           sys.executable, os.path.abspath(__file__), "--traffic-probe", "--scene", a.scene, "--width", str(a.width), "--height", str(a.height),
           "--spp-per-step", str(a.spp_per_step), "--traversal", str(a.traversal), "--kernel", a.kernel]
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    except Exception as e:  # noqa: BLE001
        return None, f"ncu side-run failed: {e}"
    tot, unit_mul = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    found = 0
    for ln in (r.stdout + "\n" + r.stderr).splitlines():
        m = re.search(r'"dram__bytes_(?:read|write)\.sum","(\w+)","([0-9.,]+)"', ln)
        if m:
            tot += float(m.group(2).replace(",", "")) * unit_mul.get(m.group(1), 1.0)
            found += 1
    if found < 2:
        tail = ((r.stderr or r.stdout).strip().splitlines() or ["?"])[-1][:160]
        return None, "ncu side-run produced no dram counters: " + tail
    return tot, "dram__bytes_read.sum + dram__bytes_write.sum of one launch, measured by an ncu side-run of this command inside this bench run"


def run_traffic_probe(a):
    import torch
    torch.cuda.set_device(0)
    sc = make_scene(a)
    b = Bench(a, sc, a.spp_per_step, options=kernel_options(a))
    b.job.render_range(0, a.spp_per_step)
    torch.cuda.synchronize()
    b.close()


def kernel_options(a):
    from aurora_rendering_engine_b200 import capi
    o = dict({"lean": {capi.OPT_BAKED_KERNEL: 0}, "generic": {capi.OPT_BAKED_KERNEL: 0, capi.OPT_LEAN_KERNEL: 0},
              "baked-packed": {capi.OPT_BAKED_PACKED: 1}}.get(a.kernel, {}))
    if a.l2_persist:
        o[capi.OPT_L2_PERSIST_NODES] = a.l2_persist
    if a.no_quant:
        o[capi.OPT_QUANTIZED_NODES] = 0
    if a.baked_min_blocks:
        o[capi.OPT_BAKED_MIN_BLOCKS] = a.baked_min_blocks
    if a.traversal == 4:  # ARE_TRAVERSAL_BVH4 needs the 4-wide collapse of the host-built tree
        o[capi.OPT_BUILD_BVH4] = 1
    return o


def measure_other_config(a, name, S, peaks, steps=3, warmup=3):
    """One of the non-headline BASELINE configurations, measured like the headline in short steps."""
    sc = make_scene(a, name)
    builder = 1 if name == "stress" else None  # 1 M primitives: the device LBVH builder (0.5 ms) instead of the host SAH one
    b = Bench(a, sc, S, builder=builder)
    try:
        t = b.timed_steps(steps, warmup, None)
        st = b.counted_chunk(warmup)
        kernel_ms = sum(t["step_ms"]) / len(t["step_ms"])
        entry = {"workload": workload_name(sc, S), "value": t["value"], "unit": UNIT, "ms_per_step": t["total_ms"] / steps, "steps": steps, "warmup": warmup,
                 "mrays_per_s": t["value"] * st.rays / max(1, st.samples), "rays_per_sample": st.rays / max(1, st.samples),
                 "commit_s": b.commit_s, "bvh_builder": "device LBVH" if builder == 1 else "host SAH",
                 "roofline": b.roofline(st, kernel_ms, peaks)}
        if not a.no_e2e:
            # the scene is committed once per job (a 1 M-primitive commit is host work of ~0.15 s, not a per-step input)
            entry["e2e"] = b.e2e(2, 1, recommit=False)
            entry["e2e"]["note"] = "scene committed once per job (commit_s beside); every step downloads the frame"
    finally:
        b.close()
    if b.world == 1 and b.rank == 0 and not a.no_cpu_baseline:
        entry["cpu_baseline"] = cpu_baseline(sc, min(a.cpu_seconds, 6.0))
    return entry


def run_b200(a):
    import torch
    import torch.distributed as dist
    from aurora_rendering_engine_b200 import engine

    rank, local_rank, world = engine.dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    sc = make_scene(a)
    S = a.spp_per_step
    b = Bench(a, sc, S, options=kernel_options(a), builder=a.builder if a.builder >= 0 else None)
    W, H = sc.width, sc.height
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t = b.timed_steps(a.steps, a.warmup, flush)
    clk = clocks.stop() if rank == 0 else None
    value, total_ms, step_ms = t["value"], t["total_ms"], t["step_ms"]
    st = b.counted_chunk(a.warmup)
    rays_per_sample = st.rays / max(1, st.samples)
    avg_kernel_ms = sum(step_ms) / max(1, len(step_ms))

    e2e = None if a.no_e2e else b.e2e(max(2, min(a.steps, 4)), 2)
    check = reduce_check(b) if world > 1 else None
    strong = None if (a.no_strong or sc.name != "cornell_box") else strong_job(b, a.job_spp)

    peaks = None
    if rank == 0:
        pk = b.ctx.measure_fp32_peak()
        l2 = b.ctx.measure_l2_peak()
        peaks = {"fp32_tflops": pk["tflops"], "sm_count": pk["sm_count"], "l2_gbs": l2["gb_per_s"], "l2_buffer_bytes": l2["buffer_bytes"]}
    roof = b.roofline(st, avg_kernel_ms, peaks) if rank == 0 else None
    b.close()
    del flush
    torch.cuda.empty_cache()

    configs = []
    if not a.no_configs and sc.name == "cornell_box":
        peaks_all = peaks
        if world > 1:  # every rank needs the peaks (values only matter on rank 0)
            obj = [peaks]
            dist.broadcast_object_list(obj, src=0)
            peaks_all = obj[0]
        for name, s_cfg in OTHER_CONFIGS:
            try:
                configs.append(measure_other_config(a, name, s_cfg, peaks_all))
            except Exception as e:  # noqa: BLE001 — a failing side config must not lose the headline line
                configs.append({"workload": name, "error": str(e)[:300]})

    if rank == 0:
        nominal = peaks["sm_count"] * 128 * 2 * (clk["sm_max_mhz"] or 1965.0) * 1e6 / 1e12 if clk else None
        # what the reference's algorithm (every ray tests every primitive of the ObjectSet, 51 FLOP per Moeller-Trumbore
        # test, 28 per sphere, 45 per quad) would have executed for the same rays — the kernel does less (fusion, boxes, BVH)
        ref_flops = st.rays * (F_TRI * len(sc.tris) + F_QUAD * len(sc.quads) + F_SPH * len(sc.spheres)) + F_SCATTER * max(0, st.rays - st.samples) + F_PRIMARY * st.samples
        traffic, traffic_src = (None, "skipped") if (a.no_traffic or world > 1) else traffic_probe(a)
        try:
            hbm_peak, hbm_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst)"
        except Exception:
            hbm_peak, hbm_src = 7700.0, "fallback: nominal HBM3e figure of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
        hbm_gbs = (2.0 * W * H * 12) / (avg_kernel_ms * 1e-3) / 1e9
        roof_hbm = {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak, "traffic": traffic,
                    "peak_source": hbm_src, "note": "not the binding resource: shown so that the FP32-issue bound in `roofline` is a measured statement"}
        roof.update({"traffic": traffic, "traffic_source": traffic_src, "reference_algorithm_flops_per_launch": ref_flops,
                     "peak_source": "measured on this GPU by are_cuda_measure_fp32_peak (register-resident FFMA loop); MEASURED_PEAKS.json has no FP32 figure",
                     "nominal_peak": nominal, "hbm_algorithmic_gbs": hbm_gbs, "l2_peak_gbs": peaks["l2_gbs"],
                     "box_unit": f"a parallelepiped (three slab pairs) test is counted as {F_BOX:.0f} FLOP; SURVEY §8d has no figure for it"})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": total_ms / max(1, a.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(a, sc, world), "mrays_per_s": value * rays_per_sample,
                "rays_per_sample": rays_per_sample, "reduce_ms": t["reduce_ms"], "clocks": clk, "e2e": e2e, "gpu_launches": a.steps,
                "roofline": roof, "roofline_hbm": roof_hbm}
        if check is not None:
            line["reduce_check"] = check
        if strong is not None:
            line["strong_job"] = strong
        if configs:
            line["configs"] = configs
        if not a.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(sc, a.cpu_seconds)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    a = parse()
    # Libraries chat on stdout at the C level (NCCL prints "NCCL version ..." on communicator creation): keep the
    # process's stdout for the JSON line only and send everything else to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.traffic_probe:
        run_traffic_probe(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
