#!/usr/bin/env python
"""Write the CUBIN of the scene-specialised kernel the library compiled for a scene ON THIS BOX (are_cuda_get_baked_cubin),
so that `nvdisasm -g` / tools/ncu_cubin_lines.py can map an ncu capture taken here to source lines.
  python tools/dump_baked_cubin.py cornell_box gpurun_out/baked.cubin"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aurora_rendering_engine_b200 import capi, scenes  # noqa: E402

sc = scenes.by_name(sys.argv[1])
with capi.Context(0) as ctx:
    sc.feed(ctx)
    ctx.commit()
    info = ctx.commit_info()
    open(sys.argv[2], "wb").write(ctx.baked_cubin())
    print(f"baked={info.baked} compile {info.bake_compile_ms:.0f} ms -> {sys.argv[2]}")
