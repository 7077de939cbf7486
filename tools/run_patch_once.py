#!/usr/bin/env python
"""tools/run_patch_once.py — one patch-renderer call at 4K (for `ncu` launch lists / captures of k_patch_nodes, k_patch_camera)."""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aurora_rendering_engine_b200 import capi, scenes  # noqa: E402

ps = dataclasses.replace(scenes.patch_rt10(width=3840, height=2160), max_tex_res=1024)
with capi.Context(0) as ctx:
    for _ in range(3):
        rgb, rgb8, st = ctx.patch_render(ps)
    print(st.as_dict())
