"""aurora_rendering_engine_b200 — B200-native path-tracing core behind the Aurora Rendering Engine API.

The product is ``lib/libare_b200.so`` (CUDA kernels for sm_100a + the C ABI of ``include/are_cuda.h``), built
from ``csrc/``.  This Python package is host-side plumbing over that ABI:

* ``capi``    ctypes binding, one method per ``are_cuda_*`` entry point
* ``scenes``  synthetic scene generators for the BASELINE.json configurations
* ``engine``  render-job driver: sample-range sharding across ranks + NCCL sum-reduce of the accumulators

There is no CPU implementation anywhere in this package: without the shared library or without a CUDA device
every compute call raises.
"""
from . import capi, scenes  # noqa: F401

__all__ = ["capi", "scenes"]
