"""CPU oracle for the LGG09 / GLGM06 hot path.  TEST INFRASTRUCTURE ONLY.

A NumPy float64 restatement of the reference's algorithm (ReachabilityAnalysis.jl v0.32.0,
`src/Algorithms/LGG09`, `src/Algorithms/GLGM06`) plus the LazySets.jl rules those loops call.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the checker / CPU baseline.  The product package
(`reachabilityanalysis.jl_b200`) never imports it and has no CPU fallback.

Pinning status:
  * the reference itself (Julia) cannot run in the build container -- no `julia` binary, and its
    arithmetic lives in LazySets.jl (compat "7", Project.toml:34), which is not vendored under
    /root/reference.  The LazySets rules restated here are the published algorithms.
  * pinned by the reference's own known-answer tests: test/flowpipes/flowpipes.jl:43-44
    (GLGM06 homogeneous, all printed digits), test/models/generated/Building.jl:44-47,52-55
    (LGG09 inhomogeneous, Forward and NoBloating), the generator-count pins of
    test/algorithms/GLGM06.jl:61,102-104 and test/continuous/solve.jl:15-16, the direction order
    of test/algorithms/LGG09.jl:52-55 -- see tests/test_oracle_kat.py.
  * GIR05 generator *selection* and the ISS / Heat3D numerics have no golden vector in the
    reference's tests: for those the oracle is "parity unpinned" beyond the rules above.
"""
