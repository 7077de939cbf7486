"""Host logic of the multi-GPU job driver: sample sharding, chunking, and the N>1 reduce path on gloo (CPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aurora_rendering_engine_b200 import engine, scenes


@pytest.mark.parametrize("spp,world", [(16384, 1), (16384, 8), (1000, 3), (7, 8), (0, 4), (500, 4)])
def test_shard_samples_cover_disjoint(spp, world):
    seen = []
    for r in range(world):
        b, c = engine.shard_samples(spp, world, r)
        assert c >= 0 and abs(c - spp / world) < 1
        seen += list(range(b, b + c))
    assert seen == list(range(spp))


def test_shard_samples_rejects_bad_input():
    for args in [(10, 0, 0), (10, 2, 2), (10, 2, -1), (-1, 2, 0)]:
        with pytest.raises(ValueError):
            engine.shard_samples(*args)


def test_chunk_ranges():
    assert engine.chunk_ranges(5, 10, 4) == [(5, 4), (9, 4), (13, 2)]
    assert engine.chunk_ranges(0, 0, 4) == []
    with pytest.raises(ValueError):
        engine.chunk_ranges(0, 4, 0)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert engine.dist_env() == (rank, rank, world)
    # each rank "renders" its own sample shard: here a deterministic stand-in per global sample index
    spp, H, W = 10, 4, 5
    b, c = engine.shard_samples(spp, world, rank)
    acc = torch.zeros((H, W, 3), dtype=torch.float32)
    for s in range(b, b + c):
        acc += torch.full((H, W, 3), float(s + 1))
    engine.reduce_sum_to_root(acc, world)
    q.put((rank, acc.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_on_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.all(res[0] == sum(range(1, 11)))  # root holds the sum over ALL samples 1..10


def test_scene_generators_shapes():
    s = scenes.rt_cornell()
    assert s.num_prims == 34 and len(s.tris) == 34  # experiments/rt.cpp:153-185: 5 quads + 2 boxes = 17 quads
    s = scenes.cornell_box()
    assert s.num_prims == 36 and (s.width, s.height, s.spp, s.max_depth) == (2048, 2048, 16384, 50)
    assert scenes.cornell_box(as_quads=True).num_prims == 6 + 24
    s = scenes.rtiow_final()
    assert 470 <= s.num_prims <= 490 and (s.width, s.height, s.spp) == (1200, 675, 500)
    s = scenes.textured()
    assert (s.width, s.height, s.spp) == (1920, 1080, 1024) and s.textures[2][2].shape == (512, 1024, 3)
    s = scenes.stress(n_prims=2000)
    assert s.num_prims == 2000 and len(s.spheres) == 1000 and len(s.tris) == 1000
    img = scenes.synthetic_image(8, 4)
    assert img[3, 5].tolist() == [5, 3, 5 ^ 3]


def test_scene_feeds_the_cpu_twin(oracle):
    for sc in (scenes.rt_cornell(), scenes.cornell_box(), scenes.textured(width=8, height=8), scenes.stress(n_prims=100)):
        osc = sc.feed(oracle.scene())
        assert osc.num_primitives() == sc.num_prims


def test_scene_compiler_fuses_parallelograms_and_boxes(lib):
    """Host-only probe of the scene compiler: the 36-triangle Cornell box becomes 18 parallelograms, of which 17 are
    faces of three parallelepipeds (the room with its front open, the two boxes); the light stays a quad."""
    import numpy as np
    from aurora_rendering_engine_b200 import capi

    def tris(sc):
        return (np.stack([t[0] for t in sc.tris]), np.stack([t[1] for t in sc.tris]), np.stack([t[2] for t in sc.tris]))

    r = capi.compile_probe(*tris(scenes.cornell_box()))
    assert (r["fused_pairs"], r["boxes"], r["brute_quads"], r["brute_tris"], r["brute_boxes"]) == (18, 3, 1, 0, 3)
    assert r["hot_slots"] == 3 * 2 + 1 and r["bvh_nodes"] == 3  # 4 hot primitives, one per leaf -> 3 inner nodes
    r = capi.compile_probe(*tris(scenes.rt_cornell()))
    assert (r["fused_pairs"], r["boxes"], r["brute_quads"]) == (17, 3, 0)
    # random triangles: nothing to fuse, a real hierarchy
    r = capi.compile_probe(*tris(scenes.stress(n_prims=2000)))
    assert r["fused_pairs"] == 0 and r["boxes"] == 0 and r["hot_slots"] == 1000 and r["bvh_nodes"] == 999 and 8 <= r["bvh_depth"] <= 40
    # a lone parallelogram pair and an open book (two quads sharing an edge) are not boxes
    Q = np.array([[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0]], float)
    u = np.array([[1, 0, 0], [1, 1, 0], [1, 0, 0], [1, 0, 1]], float)
    v = np.array([[1, 1, 0], [0, 1, 0], [1, 0, 1], [0, 0, 1]], float)
    r = capi.compile_probe(Q, u, v)
    assert r["fused_pairs"] == 2 and r["boxes"] == 0 and r["brute_quads"] == 2


def test_lean_form_and_device_builder_input(lib):
    """Host-only probe of the two derived forms of the compiled scene: the lean brute-force form (what the lean render
    kernel reads) and the input of the device BVH builder (conservative fp32 boxes, records in item order)."""
    import numpy as np
    from aurora_rendering_engine_b200 import capi

    def tris(sc):
        return (np.stack([t[0] for t in sc.tris]), np.stack([t[1] for t in sc.tris]), np.stack([t[2] for t in sc.tris]))

    # Cornell box: 3 boxes (the room first: it is the open one) + the light quad -> 3 * 6 + 1 shading records
    r = capi.compile_probe_forms(*tris(scenes.cornell_box()))
    assert (r["lean_ok"], r["lean_records"], r["lean_open_boxes"], r["brute_boxes"]) == (1, 19, 1, 3)
    assert r["lbvh_items"] == 4 and r["lbvh_slots"] == 7 and r["host_bvh_nodes"] == 0 and r["lbvh_conservative"] == r["lbvh_items"]
    # more than four loose triangles: no lean form; every item box contains its fp64 vertices after rounding to fp32
    sc = scenes.stress(n_prims=4000)
    r = capi.compile_probe_forms(*tris(sc))
    assert r["lean_ok"] == 0 and r["lean_records"] == 0
    assert r["lbvh_items"] == r["lbvh_slots"] == len(sc.tris) and r["lbvh_conservative"] == r["lbvh_items"]
    # vertices far from the origin with awkward fractions: rounding to fp32 must go outwards
    rng = np.random.RandomState(11)
    Q = rng.uniform(-1, 1, (500, 3)) * 1e4 + 1.0 / 3.0
    u, v = rng.uniform(-1, 1, (500, 3)) / 7.0, rng.uniform(-1, 1, (500, 3)) / 7.0
    r = capi.compile_probe_forms(Q, u, v)
    assert r["lbvh_conservative"] == r["lbvh_items"] == 500
    # four loose triangles still fit the lean form
    r = capi.compile_probe_forms(Q[:4] * 1e-4, u[:4], v[:4])
    assert (r["lean_ok"], r["lean_records"], r["lean_open_boxes"]) == (1, 4, 0)


def test_parallel_bvh_build_is_independent_of_thread_count(lib, monkeypatch):
    """The host BVH builder hands subtrees to threads that write straight into their (pre-computable) place of the
    depth-first arrays: the compiled hierarchy must be byte-identical for 1 thread and for many."""
    import numpy as np
    from aurora_rendering_engine_b200 import capi
    sc = scenes.stress(n_prims=120_000)
    Q, u, v = (np.stack([t[k] for t in sc.tris]) for k in range(3))
    digests = []
    for threads in (1, 3, 16):
        capi.set_build_threads(threads)
        r = capi.compile_probe(Q, u, v)
        assert r["bvh_nodes"] == len(Q) - 1
        digests.append(r["digest"])
    capi.set_build_threads(0)
    assert digests[0] == digests[1] == digests[2] and digests[0] != 0
    # a mesh with shared edges: the sorted edge table finds the same parallelograms as before (18 pairs, 3 boxes)
    box = scenes.cornell_box()
    r = capi.compile_probe(*(np.stack([t[k] for t in box.tris]) for k in range(3)))
    assert (r["fused_pairs"], r["boxes"]) == (18, 3)


def test_resume_and_takeover_bookkeeping():
    """Checkpoint / failed-rank recovery is pure interval arithmetic on GLOBAL sample indices."""
    from aurora_rendering_engine_b200 import engine
    assert engine.merge_ranges([(8, 4), (0, 4), (4, 4), (20, 2), (21, 3)]) == [(0, 12), (20, 4)]
    assert engine.missing_ranges(0, 16, []) == [(0, 16)]
    assert engine.missing_ranges(0, 16, [(0, 16)]) == []
    assert engine.missing_ranges(4, 10, [(0, 6), (8, 2), (30, 5)]) == [(6, 2), (10, 4)]
    assert engine.missing_ranges(0, 8, [(2, 2), (2, 3)]) == [(0, 2), (5, 3)]
    assert not engine.overlapping([(0, 4), (4, 4)]) and engine.overlapping([(0, 5), (4, 4)])
    # 4 ranks render 1000 spp; rank 2 dies after 100 of its samples: what is missing is exactly the rest of its shard,
    # and re-sharding it over the 3 survivors covers it without overlap
    spp, world = 1000, 4
    shards = [engine.shard_samples(spp, world, r) for r in range(world)]
    done = [s for r, s in enumerate(shards) if r != 2] + [(shards[2][0], 100)]
    miss = engine.missing_ranges(0, spp, done)
    assert miss == [(shards[2][0] + 100, shards[2][1] - 100)]
    redo = []
    for mb, mc in miss:
        for r in range(3):
            b, c = engine.shard_samples(mc, 3, r)
            redo.append((mb + b, c))
    assert not engine.overlapping(done + redo) and engine.merge_ranges(done + redo) == [(0, spp)]
