"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs ``/root/reference``; the GPU box does not have it):

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by ``radar_point_cloud_tracking_b200.synthetic`` (bit-identical
on CPU and GPU), so only the reference's OUTPUTS are stored, as small ``.npz`` files:

* ``pipeline_small.npz``  — T4 functions on a 14-frame x 3-gain x 64-spoke CSV tree, called the
  way ``run_pipeline`` calls them (T4:941-977): per-frame points/gains, occupancy grids, edges,
  land mask, filtered frames, ST-DBSCAN cluster records, and the ``clusters.csv`` of a real
  ``run_pipeline`` run on the same tree.
* ``sweeps.npz``          — ``load_radar_csv`` numeric part (T4:200-232, ``pd.read_csv`` patched to
  return the in-memory sweep) at several threshold/stride settings; digests for large outputs.
* ``package.npz``         — ``radar_pipeline`` ``polar_to_cartesian`` / ``sweep_to_point_cloud`` /
  ``st_dbscan`` (3-D coords, gain-index times — BASELINE config 2 style).
* ``fuse_max.npz``        — T5 ``fuse_gains_max`` (T5:222-273).
* ``stdbscan_random.npz`` — 24 random ST-DBSCAN problems labelled by T3's ``st_dbscan`` (T3:101-136).
* ``trig_lut.npz``        — numpy float32 cos/sin of every angle code 0..8196 on this host (digest
  only; records that the host trig path is the contract, SURVEY §7 hard part 1).
"""
from __future__ import annotations

import contextlib
import hashlib
import importlib.util
import io
import sys
import tempfile
from pathlib import Path

import numpy as np
import pandas as pd

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF / "radar-pipeline" / "src"))

from radar_point_cloud_tracking_b200 import synthetic as syn  # noqa: E402


def load_script(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, REF / rel)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sweep_dataframe(spec: syn.SweepSpec, echo: np.ndarray, gain: int) -> pd.DataFrame:
    cols = ["Status", "Scale", "Range", "Gain", "Angle"] + [f"Echo_{i}" for i in range(spec.bins)]
    meta = np.column_stack([np.ones(spec.spokes), spec.scale().astype(np.float64),
                            np.full(spec.spokes, 3.0), np.full(spec.spokes, float(gain)),
                            spec.angle_units().astype(np.float64)])
    return pd.DataFrame(np.concatenate([meta, echo.astype(np.float64)], axis=1), columns=cols)


sys.path.insert(0, str(HERE))
from specs import CLUSTER3D_SPEC, DBSCAN_CASES, PIPE_SPEC, SWEEP_CASES, SWEEP_SPEC  # noqa: E402


def make_pipeline_small(T4):
    spec = syn.SweepSpec(**PIPE_SPEC)
    echo = syn.synth_echo(spec)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        root = Path(tmp) / "data"
        frame_files = syn.write_csv_tree(spec, root, echo)
        # discovery/grouping of the reference must reproduce the same frames
        grouped = T4.group_files_by_frame(T4.discover_files(root))
        assert [sorted(g) for g in grouped] == [sorted(f) for f in frame_files]
        frames = [T4.build_frame(ff, i) for i, ff in enumerate(grouped)]
        assert all(f is not None for f in frames)
        out["frame_offsets"] = np.cumsum([0] + [f.num_points for f in frames]).astype(np.int64)
        out["points"] = np.concatenate([f.points for f in frames])
        out["gains"] = np.concatenate([f.gains for f in frames])
        count, isum, (xe, ye) = T4.build_occupancy_grid(frames, T4.LAND_GRID_RESOLUTION)
        land = T4.identify_land_cells(count, isum, len(frames))
        out.update(count=count, isum=isum, x_edges=xe, y_edges=ye, land=land)
        filt = [T4.filter_land_from_frame(f, land, (xe, ye)) for f in frames]
        out["filt_offsets"] = np.cumsum([0] + [f.num_points for f in filt]).astype(np.int64)
        out["filt_points"] = np.concatenate([f.points for f in filt])
        out["filt_gains"] = np.concatenate([f.gains for f in filt])
        # labels through the flat twin (same body) + cluster records through T4's own st_dbscan
        for tag, eps_s, eps_t, ms in DBSCAN_CASES:
            clusters = T4.st_dbscan(filt, eps_s, eps_t, ms)
            rec = []
            for fid, cl in clusters.items():
                for c in cl:
                    rec.append((fid, c.cluster_id, c.num_points, c.centroid[0], c.centroid[1],
                                c.mean_intensity))
            out[f"clusters_{tag}"] = np.array(sorted(rec), dtype=np.float64).reshape(-1, 6)
        # the real CLI path end to end (writes clusters.csv etc.)
        outdir = Path(tmp) / "out"
        with contextlib.redirect_stdout(io.StringIO()):
            T4.run_pipeline(root, outdir, visualize=False)
        cdf = pd.read_csv(outdir / "clusters.csv")
        out["clusters_csv"] = cdf.to_numpy(dtype=np.float64)
        out["clusters_csv_columns"] = np.array(list(cdf.columns))
        out["tracked_objects_csv"] = np.array((outdir / "tracked_objects.csv").read_text())
    np.savez_compressed(HERE / "pipeline_small.npz", **out)
    print("pipeline_small:", {k: getattr(v, "shape", None) for k, v in out.items()})
    return spec, echo, out


def make_sweeps(T4):
    spec = syn.SweepSpec(**SWEEP_SPEC)
    echo = syn.synth_sweep(spec, 0, 2)
    df = sweep_dataframe(spec, echo, 75)
    orig = T4.pd.read_csv
    out = {}
    try:
        T4.pd.read_csv = lambda *a, **k: df
        for tag, thr, stride in SWEEP_CASES:
            T4.INTENSITY_THRESHOLD, T4.POINT_STRIDE = thr, stride
            x, y, z, gain = T4.load_radar_csv(Path("unused.csv"))
            assert gain == 75 or len(x) == 0
            out[f"{tag}_n"] = np.int64(len(x))
            out[f"{tag}_digest"] = np.array(digest(x) + digest(y) + digest(z))
            if len(x) <= 20000:
                out[f"{tag}_x"], out[f"{tag}_y"], out[f"{tag}_z"] = x, y, z
    finally:
        T4.pd.read_csv = orig
        T4.INTENSITY_THRESHOLD, T4.POINT_STRIDE = 10.0, 4
    np.savez_compressed(HERE / "sweeps.npz", **out)
    print("sweeps:", {k: (int(v) if k.endswith("_n") else None) for k, v in out.items() if k.endswith("_n")})


def make_package():
    from radar_pipeline.config import ProcessingConfig
    from radar_pipeline.core.loaders import RadarSweep
    from radar_pipeline.core.transforms import polar_to_cartesian, sweep_to_point_cloud
    from radar_pipeline.processors.clustering import st_dbscan

    spec = syn.SweepSpec(**SWEEP_SPEC)
    out = {}
    angles = np.deg2rad(spec.angle_units().astype(np.float32) * (360.0 / 8196.0))
    ranges = (spec.scale()[:, None] / spec.bins) * np.arange(spec.bins, dtype=np.float32)
    x, y = polar_to_cartesian(angles, ranges)
    out["p2c_digest"] = np.array(digest(x) + digest(y))
    echo = syn.synth_sweep(spec, 0, 1)
    sweep = RadarSweep(angles_rad=angles, ranges=ranges, intensities=echo, scale=spec.scale(),
                       gain=50, source_path=None)
    for tag, thr, stride in (("default", 0.0, 16), ("t10_s4", 10.0, 4)):
        pc = sweep_to_point_cloud(sweep, ProcessingConfig(intensity_threshold=thr, point_stride=stride))
        out[f"s2pc_{tag}_n"] = np.int64(len(pc.x))
        out[f"s2pc_{tag}_digest"] = np.array(digest(pc.x) + digest(pc.y) + digest(pc.z))
    # config-2 style clustering: 3-D coords (x, y, intensity), "time" = gain index, eps 5/1/10
    spec3 = syn.SweepSpec(**CLUSTER3D_SPEC)
    pts, tms = [], []
    for gi in range(3):
        e = syn.synth_sweep(spec3, 0, gi)
        sw = RadarSweep(angles_rad=np.deg2rad(spec3.angle_units().astype(np.float32) * (360.0 / 8196.0)),
                        ranges=(spec3.scale()[:, None] / spec3.bins) * np.arange(spec3.bins, dtype=np.float32),
                        intensities=e, scale=spec3.scale(), gain=spec3.gains[gi], source_path=None)
        pc = sweep_to_point_cloud(sw, ProcessingConfig(intensity_threshold=10.0, point_stride=2))
        pts.append(np.column_stack((pc.x, pc.y, pc.z)))
        tms.append(np.full(len(pc.x), gi, dtype=np.float32))
    coords, times = np.concatenate(pts), np.concatenate(tms)
    out["cluster3d_n"] = np.int64(len(coords))
    out["cluster3d_labels"] = st_dbscan(coords, times, 5.0, 1.0, 10)
    np.savez_compressed(HERE / "package.npz", **out)
    print("package: cluster3d n =", len(coords), "clusters =", int(out["cluster3d_labels"].max()) + 1)


def make_fuse_max(T5, spec, echo):
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        sub = syn.SweepSpec(**{**PIPE_SPEC, "frames": 1})
        files = syn.write_csv_tree(sub, Path(tmp), echo[:1])
        for tag, res in (("r1", 1.0), ("r2p5", 2.5)):
            ox, oy, oi = T5.fuse_gains_max(files[0], grid_resolution=res)
            out[f"{tag}_x"], out[f"{tag}_y"], out[f"{tag}_i"] = ox, oy, oi
        ax, ay, ai, ag = T5.fuse_gains_absolute(files[0])
        out["abs_digest"] = np.array(digest(ax) + digest(ay) + digest(ai) + digest(ag))
        out["abs_n"] = np.int64(len(ax))
    np.savez_compressed(HERE / "fuse_max.npz", **out)
    print("fuse_max:", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


def make_stdbscan_random(T3):
    rng = np.random.default_rng(424242)
    out = {}
    for k in range(24):
        n = int(rng.integers(40, 1600))
        dim = 2 if k % 3 else 3
        span = float(rng.uniform(30, 150))
        coords = (rng.random((n, dim)) * span).astype(np.float32)
        m = n // 2
        centres = coords[rng.integers(0, 6, m)]
        coords[:m] = (centres + rng.normal(0, 2.5, (m, dim))).astype(np.float32)
        if k % 4 == 0:
            times = (rng.random(n) * 6).astype(np.float32)          # fractional times
        else:
            times = rng.integers(0, 9, n).astype(np.float32)         # frame ids
        eps_s = float(np.round(rng.uniform(2, 11), 3))
        eps_t = float(rng.choice([0.0, 1.0, 2.0, 1.5, 0.7, 3.0]))
        ms = int(rng.integers(2, 14))
        if k == 5:
            ms = 1                                                    # every point is a core
        labels = T3.st_dbscan(coords, times, eps_s, eps_t, ms)
        out[f"c{k}_coords"], out[f"c{k}_times"] = coords, times
        out[f"c{k}_params"] = np.array([eps_s, eps_t, ms], dtype=np.float64)
        out[f"c{k}_labels"] = labels
    np.savez_compressed(HERE / "stdbscan_random.npz", **out)
    print("stdbscan_random: 24 cases")


def make_trig():
    codes = np.arange(0, 8197, dtype=np.float32)
    rad = np.deg2rad(codes * (360.0 / 8196.0))
    np.savez_compressed(HERE / "trig_lut.npz",
                        cos_digest=np.array(digest(np.cos(rad[:, None]))),
                        sin_digest=np.array(digest(np.sin(rad[:, None]))),
                        numpy_version=np.array(np.__version__))


def main():
    T4 = load_script("ref_T4", "PointCloudWork/4_temporal_object_tracker.py")
    T3 = load_script("ref_T3", "PointCloudWork/3_stdbscan_point_clouds.py")
    T5 = load_script("ref_T5", "PointCloudWork/5_gain_fusion_ply_builder.py")
    spec, echo, _ = make_pipeline_small(T4)
    make_sweeps(T4)
    make_package()
    make_fuse_max(T5, spec, echo)
    make_stdbscan_random(T3)
    make_trig()


if __name__ == "__main__":
    main()
