"""Per-frame cluster records computed on the device (rb_cluster_records; SURVEY section 8 f rank 1) against the
reference's own host loop (T4:511-534: boolean masks, ``np.mean`` centroid, ``np.mean`` intensity) on the same labels:
same frames, same cluster order inside a frame (Python's set iteration order), identical member arrays, and centroids /
mean intensities equal BIT FOR BIT on non-integer data (numpy's float32 reduction orders are reproduced, not
approximated). The golden cluster records of the unmodified reference are covered through ``tracker.st_dbscan`` in
tests/test_gpu_parity.py and tests/test_gpu_reference_install.py."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(0)
    from radar_point_cloud_tracking_b200 import device as dev
    return dev


def _host_records(pts, labels, off, frame_ids):
    """T4:511-534 verbatim in spirit: per frame, for lbl in set(frame_labels) minus {-1}."""
    out = {}
    for i, fid in enumerate(frame_ids):
        lab = labels[off[i]:off[i + 1]]
        xy, inten = pts[off[i]:off[i + 1], :2], pts[off[i]:off[i + 1], 2]
        ids = set(lab)
        ids.discard(-1)
        for c in ids:
            m = lab == c
            out.setdefault(int(fid), []).append((int(c), xy[m], inten[m], np.mean(xy[m], axis=0), float(np.mean(inten[m]))))
    return out


def _compare(gpu, pts, labels, sizes, n_clusters, frame_ids):
    from radar_point_cloud_tracking_b200.device import PointBatch
    from radar_point_cloud_tracking_b200.tracker import Cluster
    d = torch.device("cuda:0")
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    t = torch.from_numpy(np.ascontiguousarray(pts.T)).to(d)
    batch = PointBatch(t[0].contiguous(), t[1].contiguous(), t[2].contiguous(), torch.zeros(len(pts), dtype=torch.int32, device=d),
                       torch.from_numpy(off).to(d), len(pts))
    rec = gpu.cluster_records(batch, torch.from_numpy(labels).to(d), n_clusters)
    got = gpu.clusters_from_records(rec, frame_ids, Cluster)
    want = _host_records(pts, labels, off, frame_ids)
    assert list(got) == list(want)                                          # same frames, same order
    n_seg = 0
    for fid, lst in want.items():
        g = got[fid]
        assert [c.cluster_id for c in g] == [w[0] for w in lst], fid       # set(frame_labels) iteration order
        for c, (cid, xy, inten, cen, mi) in zip(g, lst):
            assert np.array_equal(c.points, xy) and np.array_equal(c.intensities, inten)
            assert c.points.dtype == np.float32 and c.centroid.dtype == cen.dtype and c.centroid.shape == (2,)
            assert np.array_equal(c.centroid, cen), (fid, cid, c.centroid, cen)
            assert c.mean_intensity == mi
            n_seg += 1
    # the table's own mean-intensity column (what clusters.csv would print) equals numpy's, bit for bit
    keep = rec["label"] >= 0
    table = {(int(f), int(l)): float(m) for f, l, m in zip(rec["frame"][keep], rec["label"][keep], rec["mean_intensity"][keep])}
    for k, (fid, lst) in enumerate(want.items()):
        for cid, _, _, _, mi in lst:
            assert table[(list(frame_ids).index(fid), cid)] == mi
    assert int(keep.sum()) == n_seg
    return n_seg


def test_cluster_records_random_non_integer_data(gpu):
    """40 frames of 0..6000 points, 300 cluster ids with a skewed distribution (segments from 1 to several thousand
    points: one-tile and chained multi-tile partitions, the <8 / <=128 / recursive branches of numpy's pairwise sum),
    non-integer coordinates and intensities, empty frames, frames of noise only."""
    rng = np.random.default_rng(8)
    K = 300
    sizes = rng.integers(0, 6000, 40)
    sizes[[3, 17]] = 0
    pts, labels = [], []
    for i, n in enumerate(sizes):
        p = np.column_stack([rng.normal(0, 150, n), rng.normal(0, 150, n), rng.random(n) * 255]).astype(np.float32)
        if i == 5:
            lab = np.full(n, -1)
        else:
            lab = np.minimum((rng.pareto(0.6, n) * 3).astype(np.int64), K) - 1          # -1 .. K-1, heavy head, long tail
        pts.append(p)
        labels.append(lab.astype(np.int32))
    pts, labels = np.concatenate(pts), np.concatenate(labels)
    n_seg = _compare(gpu, pts, labels, sizes, K, np.arange(100, 140))
    assert n_seg > 1500


def test_cluster_records_one_huge_frame_and_many_tiny(gpu):
    """A frame of 300 k points whose biggest cluster spans ~290 tiles (the in-frame tile chain), next to 2000 frames of
    a handful of points each."""
    rng = np.random.default_rng(9)
    sizes = np.concatenate([[300_000], rng.integers(0, 8, 2000)])
    n = int(sizes.sum())
    pts = np.column_stack([rng.normal(0, 100, n), rng.normal(0, 100, n), rng.random(n) * 200 + 0.37]).astype(np.float32)
    labels = rng.integers(-1, 5, n).astype(np.int32)
    labels[:300_000][rng.random(300_000) < 0.9] = 2
    _compare(gpu, pts, labels, sizes, 5, np.arange(len(sizes)))


def test_cluster_records_of_a_pipeline_block_equal_the_host_loop(gpu):
    """A full-size 64-frame block through the pipeline: DetectionResult.clusters_by_frame (device records) against
    clusters_by_frame_host (the reference's loop on the read-back labels)."""
    from radar_point_cloud_tracking_b200 import synthetic as syn
    from radar_point_cloud_tracking_b200.pipeline import DetectionConfig, DetectionPipeline
    spec = syn.SweepSpec(seed=77, frames=64)
    pipe = DetectionPipeline(DetectionConfig(), 0)
    echo = gpu.synth_echo(spec)
    c, s, r = pipe.spoke_tables(spec.angle_units(), spec.scale(), spec.frames, spec.bins)
    res = pipe.run_device(echo, *(torch.from_numpy(t).to(echo.device) for t in (c, s, r)), frame_ids=np.arange(500, 564))
    got, want = res.clusters_by_frame(), res.clusters_by_frame_host()
    assert list(got) == list(want) and len(got) > 30
    total = 0
    for fid in want:
        assert [c.cluster_id for c in got[fid]] == [c.cluster_id for c in want[fid]]
        for a, b in zip(got[fid], want[fid]):
            assert np.array_equal(a.points, b.points) and np.array_equal(a.intensities, b.intensities)
            assert np.array_equal(a.centroid, b.centroid) and a.mean_intensity == b.mean_intensity
            total += 1
    assert total > 100
    empty = pipe.run_device(torch.zeros_like(echo), *(torch.from_numpy(t).to(echo.device) for t in (c, s, r)))
    assert empty.clusters_by_frame() == {}
