"""krisp_b200 — B200-native diagnostic-region search behind krisp_fasta / kstream.

The compute path is libkrisp_b200.so (hand-written sm_100a CUDA, C ABI in include/krisp_b200.h);
this package is the thin Python host layer that mirrors the reference's entry points.
"""
__version__ = "0.1.0"
