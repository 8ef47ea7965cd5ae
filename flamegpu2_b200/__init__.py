"""flamegpu2_b200 -- B200-native (sm_100a) implementation of FLAME GPU 2's per-step spatial hot
path (PBM build, neighbour iteration, birth/death compaction, spatial agent sort) behind the
reference's own API.  The product is the CUDA library + C++ API layer; this package is the Python
harness used by the tests, bench.py and the multi-GPU driver.  No CPU fallback exists."""

__all__ = ["_capi"]
