"""Seeds/specs shared by make_golden.py (reference side) and the tests (oracle / CUDA side)."""
PIPE_SPEC = dict(seed=20251, frames=14, spokes=64, bins=1024, clutter_p=0.01,
                 land_blobs=2, buoys=3, boats=2)
SWEEP_SPEC = dict(seed=77, frames=1, spokes=512, bins=1024, clutter_p=0.004)
CLUSTER3D_SPEC = dict(seed=5, frames=1, spokes=256, bins=1024, clutter_p=0.004)
SWEEP_CASES = (("t10_s4", 10.0, 4), ("t2_s2", 2.0, 2), ("t0_s1", 0.0, 1),
               ("t100_s3", 100.0, 3), ("t300_s4", 300.0, 4))
DBSCAN_CASES = (("default", 8.0, 2.0, 15), ("tight", 4.0, 1.0, 6))
