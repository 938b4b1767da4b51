"""B200-native alignment hot path of smritip/real-time-audio-sync.

The directory name is not a valid Python identifier (it mirrors the reference
repository's name), so import it in one of two ways:

* drop-in, like the reference's flat scripts: put this directory on ``sys.path``
  and ``import chroma, dtw, otw_eran, livenote_v2, wtw`` exactly as before;
* as a package: ``import __graft_entry__; pkg = __graft_entry__.load_package()``
  (registers it as ``rtas_b200``), then ``pkg.dtw.DTW(...)``.

Modules (same names and signatures as the reference files they replace):
``chroma``, ``dtw``, ``otw_eran``, ``livenote``, ``livenote_v2``, ``wtw``; plus
``batch`` (many pairs / streams / tracks per launch, sharded over GPUs) and
``_native`` (ctypes binding of include/afsync.h).  All arithmetic on the path runs
in libafsync.so (csrc/*.cu, sm_100a); there is no CPU fallback.
"""
__all__ = ["chroma", "dtw", "otw_eran", "livenote", "livenote_v2", "wtw", "batch", "_native"]
