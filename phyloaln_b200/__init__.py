"""phyloaln_b200: B200-native (sm_100a) search hot path for PhyloAln.

Host side (Python): HMMER3 profile loading, read ingest, ctypes binding of the CUDA library,
report thresholds and the drop-in step functions mirroring /root/reference/lib/aln.py.
Device side (csrc/): hand-written CUDA kernels behind a C ABI (include/phyloaln_b200.h).
"""
__version__ = "0.1.0"
