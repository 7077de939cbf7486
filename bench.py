#!/usr/bin/env python
"""bench.py — throughput of the path-tracing hot path on the BASELINE.json headline workload.

Workload (config.workload): BASELINE config 3 — Cornell box (5 walls + light + 2 boxes as 36 triangles, 555-unit
dimensions), 2048x2048, max depth 50.  The job's 16384 spp are rendered in chunks; ONE STEP = one chunk of
``--spp-per-step`` samples per pixel over the whole frame on every rank (weak scaling: rank r, step k renders
global samples [(k*N + r)*S, +S)), and the timed region ends with the job's NCCL sum-reduce of the accumulators.
The default chunk is 256 spp (64 launches make the job): a warp's pool is its 8x4 tile x the chunk's samples and its last
trips run on its few longest paths, so the chunk size shows — 64 / 128 / 256 / 512 spp per launch: 8175 / 8359 / 8456 / 8511
Msamples/s.

  value      Msamples/s, whole job, scene + accumulator resident in HBM, CUDA events on the launch stream,
             max over ranks.  (Mrays/s is reported beside it.)
  e2e        the same metric through the public host-buffer API: are_cuda_commit (scene H2D) + are_cuda_render
             (zero, render, W*H*3 float D2H) every step.
  roofline   the render kernel against the FP32 FMA issue peak (this path has no dense contraction and its
             working set lives in shared memory, so neither HBM nor tensor peak bounds it — DESIGN.md §5);
             achieved = counted primitive tests x SURVEY §8d FLOP figures / kernel time.
  cpu_baseline  the CPU twin (oracle/are_oracle.c, "port": the reference has no renderer for this config) on all
             host threads, on a bounded sample of the same frame.

``--impl reference`` times that CPU implementation alone (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
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
# a parallelepiped test = three slab pairs with arbitrary normals: 3 x (two dots 10 + sub 1 + rcp 1 + 2 mul + 2 add) + 4 min/max
F_BOX = 52.0


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def make_scene(a):
    from aurora_rendering_engine_b200 import scenes
    kw = dict(width=a.width, height=a.height)
    return scenes.by_name(a.scene, **kw)


BASELINE_CONFIG = {"rt_cornell": "0; reference program: 1 primary + 32 AO rays per pixel", "rtiow_final": "1; full job = 500 spp",
                   "textured": "2; full job = 1024 spp", "cornell_box": "3; full job = 16384 spp", "stress": "4; full job = 256 spp"}


def workload_config(a, sc, world):
    return {
        "workload": f"{sc.name} {a.width}x{a.height} depth {sc.max_depth} ({sc.num_prims} primitives), "
                    f"{a.spp_per_step} spp per step per GPU (BASELINE config {BASELINE_CONFIG.get(sc.name, '?')})",
        "spp_per_step": a.spp_per_step, "width": a.width, "height": a.height, "max_depth": sc.max_depth,
        "parallelism": f"sample-sharded x{world}, one NCCL sum-reduce of the accumulators at job end",
        "l2": "flushed between timed steps (256 MiB device write); per-step CUDA events summed",
    }


# ---------------------------------------------------------------------------------------------------------
# CPU implementation (oracle port) — used for cpu_baseline and for --impl reference
# ---------------------------------------------------------------------------------------------------------
class CpuTwin:
    def __init__(self, sc):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_binding import Oracle
        from aurora_rendering_engine_b200 import capi
        self.capi = capi
        self.sc = sc
        self.orc = Oracle()
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
        """Pick (spp, rows) so that one run takes about target_s: a probe of every 16th row band first."""
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
    twin = CpuTwin(sc)
    spp, rows, _ = twin.calibrate(seconds)
    dt, n, rays = twin.run(spp, rows)
    what = f"{spp} spp of " + ("the whole frame" if rows is None else f"rows {rows[0]}..{rows[1]} of the frame")
    return {"value": n / dt / 1e6, "unit": UNIT, "cores": twin.threads, "kind": "port",
            "sample": f"{what} ({n} samples, {dt:.1f} s), fp64 CPU twin oracle/are_oracle.c, pthreads over rows",
            "mrays_per_s": rays / dt / 1e6}


def run_reference(a):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    sc = make_scene(a)
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
    cfg = workload_config(a, sc, 1)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": tot_t / max(1, a.steps) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "mrays_per_s": tot_r / tot_t / 1e6,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": twin.threads, "kind": "port",
                             "sample": what + "; the reference ships no renderer for this config (SURVEY.md §0), so its CPU path is the fp64 twin in oracle/are_oracle.c"},
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
def run_b200(a):
    import torch
    import torch.distributed as dist
    from aurora_rendering_engine_b200 import capi, engine

    rank, local_rank, world = engine.dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    sc = make_scene(a)
    S = a.spp_per_step
    opts = {"lean": {capi.OPT_BAKED_KERNEL: 0}, "generic": {capi.OPT_BAKED_KERNEL: 0, capi.OPT_LEAN_KERNEL: 0},
            "baked-packed": {capi.OPT_BAKED_PACKED: 1}}.get(a.kernel, {})
    job = engine.RenderJob(sc, local_rank, a.traversal, options=opts)
    ctx = job.ctx
    W, H = sc.width, sc.height
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sample_base(step):  # global sample index of this rank's chunk in `step`
        return (step * world + rank) * S

    # --- warm-up (also yields exact counters for one chunk: paths depend only on the seed/sample indices) ---
    stats = None
    for k in range(max(a.warmup, 1) if a.warmup else 0):
        stats = job.render_range(sample_base(k), S, want_stats=True)
    if world > 1:  # warm the communicator
        dist.all_reduce(torch.zeros(1, device="cuda"))
    job.accum.zero_()
    barrier()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    red0, red1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    for k in range(a.steps):
        flush.fill_(k & 0xFF)  # L2 flush, outside the per-step event pair
        ev[k][0].record()
        job.ctx.render_device(job.cam, job.params(sample_base(a.warmup + k), S), job.accum.data_ptr())
        ev[k][1].record()
    red0.record()
    engine.reduce_sum_to_root(job.accum, world)
    red1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = sum(step_ms) + red0.elapsed_time(red1)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    samples_all = float(W) * H * S * a.steps * world
    value = samples_all / (total_ms * 1e-3) / 1e6

    # exact per-chunk counters of the first timed chunk (untimed re-run with stats), for Mrays/s + roofline
    scratch = torch.zeros_like(job.accum)
    st = ctx.render_device(job.cam, job.params(sample_base(a.warmup), S), scratch.data_ptr(), want_stats=True, count_tests=True)
    rays_per_sample = st.rays / max(1, st.samples)
    use_bvh = st.node_visits > 0
    flops = (F_TRI * st.tri_tests + F_SPH * st.sphere_tests + F_QUAD * st.quad_tests + F_BOX * st.box_tests + F_AABB * 2 * st.node_visits
             + F_SCATTER * max(0, st.rays - st.samples) + F_PRIMARY * st.samples)
    kernel_ms = step_ms[0] if step_ms else st.kernel_ms
    avg_kernel_ms = sum(step_ms) / max(1, len(step_ms))

    # --- e2e: host-buffer API, scene upload + accumulator download inside the timed region ---
    e2e = None
    if not a.no_e2e:
        host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory().numpy()
        h2d = 0
        for k in range(2):
            h2d = ctx.commit()
            ctx.render(job.cam, job.params(sample_base(k), S), out=host)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(2, min(a.steps, 4))
        e0.record()
        for k in range(n_e2e):
            h2d = ctx.commit()
            ctx.render(job.cam, job.params(sample_base(a.warmup + k), S), out=host)
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": float(W) * H * S * n_e2e * world / (float(te.item()) * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(W * H * 3 * 4), "steps": n_e2e,
               "api": "are_cuda_commit + are_cuda_render (host buffers)"}

    if rank == 0:
        peak = ctx.measure_fp32_peak()
        nominal = peak["sm_count"] * 128 * 2 * (clk["sm_max_mhz"] or 1965.0) * 1e6 / 1e12 if clk else None
        achieved = flops / (avg_kernel_ms * 1e-3) / 1e12
        # what the reference's algorithm (every ray tests every primitive of the ObjectSet, 51 FLOP per Moeller-Trumbore
        # test, 28 per sphere, 45 per quad) would have executed for the same rays — the kernel does less (fusion, boxes, BVH)
        ref_flops = st.rays * (F_TRI * len(sc.tris) + F_QUAD * len(sc.quads) + F_SPH * len(sc.spheres)) + F_SCATTER * max(0, st.rays - st.samples) + F_PRIMARY * st.samples
        traffic, ncu = None, None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this launch size, from the committed ncu capture
            prof_name = "r01_render_lean_full.json" if st.kernel_variant == capi.KERNEL_BRUTE_LEAN else "r01_render_path_full.json"
            prof = json.load(open(os.path.join(ROOT, "profiles", prof_name)))
            if prof.get("width") == W and prof.get("height") == H and prof.get("spp_per_step") == S:
                traffic = prof.get("dram_bytes")
                # the resource that actually binds this kernel (not measured live: copied from the committed capture)
                ncu = {"issue_slot_utilisation_pct": prof.get("issue_slot_utilisation_pct"),
                       "active_threads_per_instruction": prof.get("active_threads_per_instruction"),
                       "source": "profiles/%s (ncu --set full, same launch size)" % prof_name}
        except Exception:
            pass
        kernel_name = {capi.KERNEL_RT_AO: "k_render_rtao", capi.KERNEL_BRUTE: "k_render_path<brute/smem>",
                       capi.KERNEL_BRUTE_LEAN: "k_render_path<brute/smem, lean>", capi.KERNEL_BRUTE_BAKED: "k_render_baked (lean kernel, scene compiled in by NVRTC)",
                       capi.KERNEL_BVH2: "k_render_path<bvh2>",
                       capi.KERNEL_BVH2_BIG: "k_render_path<bvh2, 12 CTAs/SM>", capi.KERNEL_WIDE: "k_render_path<wide bvh>"}.get(st.kernel_variant, "?")
        # the same kernel against the HBM roofline (MEASURED_PEAKS.json, driver-written): algorithmic bytes per launch = one
        # read-modify-write of the W*H*3 fp32 accumulator; the working set of the loop lives in shared memory / registers
        try:
            hbm_peak, hbm_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst)"
        except Exception:
            hbm_peak, hbm_src = 7700.0, "fallback: nominal HBM3e figure of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
        hbm_gbs = (2.0 * W * H * 12) / (avg_kernel_ms * 1e-3) / 1e9
        roof_hbm = {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak, "traffic": traffic,
                    "peak_source": hbm_src, "note": "not the binding resource: shown so that the FP32-issue bound in `roofline` is a measured statement"}
        roof = {"bound": "fp32", "achieved": achieved, "peak": peak["tflops"], "unit": "TFLOP/s", "frac": achieved / peak["tflops"] if peak["tflops"] else None,
                "traffic": traffic, "reference_equivalent_tflops": ref_flops / (avg_kernel_ms * 1e-3) / 1e12, "ncu": ncu, "peak_source": "measured on this GPU by are_cuda_measure_fp32_peak (register-resident FFMA loop); MEASURED_PEAKS.json has no FP32 figure",
                "nominal_peak": nominal, "kernel": kernel_name, "kernel_ms": avg_kernel_ms,
                "flops_per_launch": flops, "counted": {"rays": st.rays, "tri_tests": st.tri_tests, "quad_tests": st.quad_tests,
                                                        "sphere_tests": st.sphere_tests, "box_tests": st.box_tests, "node_visits": st.node_visits},
                "hbm_algorithmic_gbs": (2.0 * W * H * 12) / (avg_kernel_ms * 1e-3) / 1e9}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": total_ms / max(1, a.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(a, sc, world), "mrays_per_s": value * rays_per_sample,
                "rays_per_sample": rays_per_sample, "reduce_ms": red0.elapsed_time(red1), "clocks": clk, "e2e": e2e, "gpu_launches": a.steps,
                "roofline": roof, "roofline_hbm": roof_hbm}
        if not a.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(sc, a.cpu_seconds)
        emit(line)
    job.close()
    if world > 1:
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
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
