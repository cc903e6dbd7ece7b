"""igd_b200 -- B200-native batch overlap search of databio/IGD.

``api``       ctypes binding of the C-ABI library (``include/igd_b200.h``): ``IGD`` (open / count / full / seqpare /
              hitsmap / search_n ...), ``create``, ``HitsReducer``
``multi``     one-process-per-GPU driver of the sharded modes (``ShardedSearch``)
``workloads`` BASELINE.json's configurations and the host-side shard planner
``synth``     synthetic BED / .igd generators for tests and benchmarks

Every search runs in the CUDA library built from ``csrc/``; there is no CPU search path in this package.
"""
