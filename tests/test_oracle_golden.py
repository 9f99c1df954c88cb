"""The oracle (oracle/) pinned against outputs of the unmodified reference (tests/golden/) and
against the reference's portable known-answer tests. CPU only."""
import numpy as np
import pytest

from oracle import numpy_oracle as O
from oracle.c_oracle import st_dbscan_c
from radar_point_cloud_tracking_b200 import synthetic as syn
from tests.common import (CLUSTER3D_SPEC, DBSCAN_CASES, SWEEP_CASES, SWEEP_SPEC, digest, golden,
                          pipe_inputs)


def oracle_frames(spec, echo, thr=10.0, stride=4):
    ang, scale = spec.angle_units(), spec.scale()
    frames = []
    for f in range(spec.frames):
        per_gain = {g: O.sweep_to_points(echo[f, gi], ang, scale, thr, stride)
                    for gi, g in enumerate(spec.gains)}
        frames.append(O.fuse_concat(per_gain))
    return frames


def test_trig_path_matches_recorded_host():
    g = golden("trig_lut")
    codes = np.arange(0, 8197, dtype=np.float32)
    rad = np.deg2rad(codes * O.ANGLE_SCALE)
    if str(g["numpy_version"]) != np.__version__:
        pytest.skip("different numpy build: trig digest is host specific")
    assert digest(np.cos(rad[:, None])) == str(g["cos_digest"])
    assert digest(np.sin(rad[:, None])) == str(g["sin_digest"])
    # contiguous and column-view evaluation agree (the host tables are computed contiguously)
    assert np.array_equal(np.cos(rad), np.cos(rad[:, None])[:, 0])
    assert np.array_equal(np.sin(rad), np.sin(rad[:, None])[:, 0])


@pytest.mark.parametrize("tag,thr,stride", SWEEP_CASES)
def test_sweep_to_points_vs_reference(tag, thr, stride):
    g = golden("sweeps")
    spec = syn.SweepSpec(**SWEEP_SPEC)
    x, y, z = O.sweep_to_points(syn.synth_sweep(spec, 0, 2), spec.angle_units(), spec.scale(), thr, stride)
    assert len(x) == int(g[f"{tag}_n"])
    assert digest(x) + digest(y) + digest(z) == str(g[f"{tag}_digest"])
    if f"{tag}_x" in g.files:
        assert np.array_equal(x, g[f"{tag}_x"]) and np.array_equal(z, g[f"{tag}_z"])


def test_frames_fusion_vs_reference():
    g = golden("pipeline_small")
    spec, echo = pipe_inputs()
    frames = oracle_frames(spec, echo)
    offs = np.cumsum([0] + [len(p) for p, _ in frames])
    assert np.array_equal(offs, g["frame_offsets"])
    assert np.array_equal(np.concatenate([p for p, _ in frames]), g["points"])
    assert np.array_equal(np.concatenate([q for _, q in frames]), g["gains"])
    assert frames[0][0].dtype == np.float32 and frames[0][1].dtype == np.int32


def test_land_filter_vs_reference():
    g = golden("pipeline_small")
    offs = g["frame_offsets"]
    pts = [g["points"][offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    count, isum, (xe, ye) = O.occupancy_grid(pts, O.LAND_GRID_RESOLUTION)
    assert xe.dtype == np.float64
    assert np.array_equal(xe, g["x_edges"]) and np.array_equal(ye, g["y_edges"])
    assert np.array_equal(count, g["count"]) and count.dtype == np.int32
    assert np.array_equal(isum, g["isum"])
    land = O.land_cells(count, isum, len(pts))
    assert np.array_equal(land, g["land"]) and land.any()
    keep = np.concatenate([O.land_keep_mask(p, land, (xe, ye)) for p in pts])
    assert np.array_equal(g["points"][keep], g["filt_points"])
    assert np.array_equal(g["gains"][keep], g["filt_gains"])


@pytest.mark.parametrize("tag,eps_s,eps_t,ms", DBSCAN_CASES)
def test_cluster_records_vs_reference(tag, eps_s, eps_t, ms):
    g = golden("pipeline_small")
    offs = g["filt_offsets"]
    frames = [(i, g["filt_points"][offs[i]:offs[i + 1]]) for i in range(len(offs) - 1)]
    for sequential in (True, False):
        labels, o2 = O.st_dbscan_frames(frames, eps_s, eps_t, ms, sequential=sequential)
        rec = []
        for fid, cl in O.clusters_per_frame(frames, labels, o2).items():
            for c in cl:
                rec.append((fid, c["cluster_id"], len(c["points"]), c["centroid"][0], c["centroid"][1],
                            float(np.mean(c["intensities"]))))
        rec = np.array(sorted(rec), dtype=np.float64).reshape(-1, 6)
        assert np.array_equal(rec, g[f"clusters_{tag}"])
    # C oracle gives the same labels
    coords = np.vstack([p[:, :2] for _, p in frames])
    fid = np.concatenate([np.full(len(p), f) for f, p in frames]).astype(np.float32)
    assert np.array_equal(st_dbscan_c(coords, fid, eps_s, eps_t, ms)[0], labels)
    if tag == "default":   # the real CLI run wrote the same rows
        csv = g["clusters_csv"]
        csv = csv[np.lexsort((csv[:, 1], csv[:, 0]))]
        # to_csv prints the float32 centroids with their shortest round-trip repr
        assert np.array_equal(csv.astype(np.float32), rec.astype(np.float32))
        assert np.array_equal(csv[:, :3], rec[:, :3])   # (pandas' CSV float parser is 1 ulp inexact in f64)


def test_random_stdbscan_vs_reference():
    g = golden("stdbscan_random")
    for k in range(24):
        coords, times = g[f"c{k}_coords"], g[f"c{k}_times"]
        eps_s, eps_t, ms = g[f"c{k}_params"]
        want = g[f"c{k}_labels"]
        assert np.array_equal(O.st_dbscan_sequential(coords, times, eps_s, eps_t, int(ms)), want), k
        assert np.array_equal(O.st_dbscan_canonical(coords, times, eps_s, eps_t, int(ms))[0], want), k
        assert np.array_equal(st_dbscan_c(coords, times, eps_s, eps_t, int(ms))[0], want), k


def test_package_functions_vs_reference():
    g = golden("package")
    spec = syn.SweepSpec(**SWEEP_SPEC)
    c, s, res = O.spoke_tables(spec.angle_units(), spec.scale(), spec.bins)
    rng = res[:, None] * np.arange(spec.bins, dtype=np.float32)
    assert digest(rng * c[:, None]) + digest(rng * s[:, None]) == str(g["p2c_digest"])
    echo = syn.synth_sweep(spec, 0, 1)
    for tag, thr, stride in (("default", 0.0, 16), ("t10_s4", 10.0, 4)):
        x, y, z = O.sweep_to_points(echo, spec.angle_units(), spec.scale(), thr, stride)
        assert len(x) == int(g[f"s2pc_{tag}_n"])
        assert digest(x) + digest(y) + digest(z) == str(g[f"s2pc_{tag}_digest"])
    spec3 = syn.SweepSpec(**CLUSTER3D_SPEC)
    pts, tms = [], []
    for gi in range(3):
        x, y, z = O.sweep_to_points(syn.synth_sweep(spec3, 0, gi), spec3.angle_units(), spec3.scale(), 10.0, 2)
        pts.append(np.column_stack((x, y, z)))
        tms.append(np.full(len(x), gi, dtype=np.float32))
    coords, times = np.concatenate(pts), np.concatenate(tms)
    assert len(coords) == int(g["cluster3d_n"])
    assert np.array_equal(O.st_dbscan_sequential(coords, times, 5.0, 1.0, 10), g["cluster3d_labels"])
    assert np.array_equal(st_dbscan_c(coords, times, 5.0, 1.0, 10)[0], g["cluster3d_labels"])


def test_fuse_max_vs_reference():
    g = golden("fuse_max")
    spec, echo = pipe_inputs()
    ang, scale = spec.angle_units(), spec.scale()
    per_gain = {gain: O.sweep_to_points(echo[0, gi], ang, scale, 5.0, 8)      # T5:57-58 defaults
                for gi, gain in enumerate(spec.gains)}
    for tag, res in (("r1", 1.0), ("r2p5", 2.5)):
        ox, oy, oi = O.fuse_max(per_gain, res)
        assert np.array_equal(ox, g[f"{tag}_x"]) and np.array_equal(oy, g[f"{tag}_y"])
        assert np.array_equal(oi, g[f"{tag}_i"]) and ox.dtype == g[f"{tag}_x"].dtype
    pts, gains = O.fuse_concat(per_gain)
    assert len(pts) == int(g["abs_n"])
    assert (digest(pts[:, 0]) + digest(pts[:, 1]) + digest(pts[:, 2]) + digest(gains)) == str(g["abs_digest"])


# ---- portable known-answer tests of the reference ------------------------------------------
def test_polar_known_answers():
    """radar-pipeline/tests/test_transforms.py:15-40 (atol 1e-6)."""
    ang = np.array([0, np.pi / 2, np.pi], dtype=np.float32)
    x, y = np.float32(1.0) * np.cos(ang), np.float32(1.0) * np.sin(ang)
    np.testing.assert_allclose(x, [1, 0, -1], atol=1e-6)
    np.testing.assert_allclose(y, [0, 1, 0], atol=1e-6)


def test_threshold_is_strict_and_stride_picks_every_nth():
    """radar-pipeline-rs/src/core/transforms.rs:471-512."""
    echo = np.array([[10.0, 50.0, 100.0], [5.0, 60.0, 200.0]], dtype=np.float32)
    _, _, z = O.sweep_to_points(echo, np.array([0, 2049]), np.array([3.0, 3.0]), 30.0, 1)
    assert sorted(z.tolist()) == [50.0, 60.0, 100.0, 200.0]
    _, _, z = O.sweep_to_points(np.full((1, 10), 100.0, np.float32), np.array([0]), np.array([10.0]), 0.0, 3)
    assert len(z) == 4
    _, _, z = O.sweep_to_points(np.full((1, 4), 30.0, np.float32), np.array([0]), np.array([10.0]), 30.0, 1)
    assert len(z) == 0


@pytest.mark.parametrize("fn", [lambda *a: O.st_dbscan_sequential(*a), lambda *a: st_dbscan_c(*a)[0]])
def test_rust_known_answer_clusters(fn):
    """radar-pipeline-rs/src/processors/clustering.rs:502-597."""
    sq = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]]
    coords = np.array(sq + [[100 + a, 100 + b, 0] for a, b, _ in sq], dtype=np.float32)
    lab = fn(coords, np.zeros(8, np.float32), 5.0, 1.0, 2)
    assert lab[0] >= 0 and len(set(lab[:4])) == 1 and len(set(lab[4:])) == 1 and lab[0] != lab[4]
    lab = fn(np.array(sq, dtype=np.float32), np.array([0, 0, 5, 5], np.float32), 5.0, 1.0, 2)
    assert lab[0] == lab[1] and lab[2] == lab[3] and lab[0] != lab[2]
    lab = fn(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [100, 100, 100]], np.float32),
             np.zeros(4, np.float32), 5.0, 1.0, 3)
    assert lab[0] >= 0 and lab[0] == lab[1] == lab[2] and lab[3] == -1
    assert len(fn(np.zeros((0, 3), np.float32), np.zeros(0, np.float32), 5.0, 1.0, 3)) == 0
    assert fn(np.zeros((1, 3), np.float32), np.zeros(1, np.float32), 5.0, 1.0, 2)[0] == -1


def test_wf_variant_oracles_vs_reference_golden():
    """PointCloudWorkF min_frames variant: the order-free numpy statement and the C oracle against labels of the
    UNMODIFIED reference function (tests/golden/make_golden_wf.py), incl. 3-D and fractional-time cases."""
    from oracle.c_oracle import st_dbscan_wf_c
    g = golden("wf_stdbscan")
    assert int(g["n_cases"]) >= 16
    for k in range(int(g["n_cases"])):
        eps_s, eps_t, ms, mf = g[f"c{k}_params"]
        c, tm, want = g[f"c{k}_coords"], g[f"c{k}_times"], g[f"c{k}_labels"]
        got_c, _ = st_dbscan_wf_c(c, tm, float(eps_s), float(eps_t), int(ms), int(mf))
        assert np.array_equal(got_c, want), k
        if k < 6:                                                 # the Python restatements are slow: a few cases
            assert np.array_equal(O.st_dbscan_wf_canonical(c, tm, float(eps_s), float(eps_t), int(ms), int(mf))[0], want), k
            assert np.array_equal(O.st_dbscan_wf_sequential(c, tm, float(eps_s), float(eps_t), int(ms), int(mf)), want), k


def test_cluster_mean_reduction_orders_are_numpys():
    """The float32 reduction orders the device cluster records reproduce (csrc/clusters.cu), pinned against numpy itself:
    np.mean over the rows of an [n, 2] array is sequential per column; np.mean of a 1-D array is numpy's pairwise sum."""
    rng = np.random.default_rng(17)
    for n in list(range(1, 40)) + [127, 128, 129, 255, 256, 257, 1000, 4097, 20001]:
        pts = (rng.random((n, 2)) * 400 - 200).astype(np.float32)
        inten = (rng.random(n) * 255).astype(np.float32)
        cen, mi = O.cluster_means_f32(pts, inten)
        assert np.array_equal(cen, np.mean(pts, axis=0)), n
        assert mi == np.mean(inten) and mi.dtype == np.float32, n
        assert O.pairwise_sum_f32(inten) == np.add.reduce(inten), n
