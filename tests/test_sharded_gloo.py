"""Host logic of the time-sharded driver on CPU: world_size 2 and 3 over gloo.

The numeric stages are played by an ORACLE-backed engine (numpy / scipy restatement, test
infrastructure), so what is under test is exactly the part of ``sharded.py`` that has no kernel:
the bounds / grid reductions, the ``floor(eps_time)``-frame halo exchange, the core-flag exchange,
the component stitching (on every rank) and the canonical global numbering. The result must equal the
single-process oracle on the concatenated recording, label for label.
"""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

from oracle import numpy_oracle as O                                   # noqa: E402
from oracle.c_oracle import st_dbscan_c                                 # noqa: E402
from radar_point_cloud_tracking_b200 import synthetic as syn           # noqa: E402
from radar_point_cloud_tracking_b200.device import PointBatch          # noqa: E402
from radar_point_cloud_tracking_b200.pipeline import DetectionConfig   # noqa: E402
from radar_point_cloud_tracking_b200.sharded import ShardedDetection, TorchEngineBase, stitch_components  # noqa: E402

SPEC = dict(seed=21, frames=12, spokes=96, bins=512, clutter_p=0.01, land_blobs=2, buoys=3, boats=3)


# ------------------------------------------------------------------------------ oracle engine (CPU)
class OraclePhases:
    """CPU twin of ``device.StDbscanPhases`` with the reference's exact predicates."""

    def __init__(self, x, y, times, eps_space, eps_time, min_samples):
        from scipy.spatial import cKDTree

        self.n = len(times)
        xy = np.column_stack([x.numpy(), y.numpy()]).astype(np.float64)
        t = times.numpy().astype(np.float32)
        eps_t = np.float32(eps_time)
        self.nbrs = []
        if self.n:
            tree = cKDTree(xy)
            cand = tree.query_ball_point(xy, r=eps_space * (1 + 1e-9) + 1e-12)
            for i, c in enumerate(cand):
                c = np.asarray(c, dtype=np.int64)
                d = xy[c] - xy[i]
                ok = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] <= eps_space * eps_space) & (np.abs(t[c] - t[i]) <= eps_t)
                self.nbrs.append(np.sort(c[ok]))
        self.min_samples = min_samples
        self.core = np.zeros(self.n, dtype=bool)

    def cores(self):
        self.core = np.array([len(nb) >= self.min_samples for nb in self.nbrs], dtype=bool)
        return torch.from_numpy(self.core.astype(np.uint8))

    def set_cores(self, core):
        self.core = core.numpy().astype(bool)

    def components(self, gidx):
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components

        g = np.arange(self.n, dtype=np.int64) if gidx is None else gidx.numpy()
        rows, cols = [], []
        for i in np.flatnonzero(self.core):
            nb = self.nbrs[i]
            nb = nb[self.core[nb]]
            rows.append(np.full(len(nb), i))
            cols.append(nb)
        key = np.full(self.n, -1, dtype=np.int64)
        if rows:
            r, c = np.concatenate(rows), np.concatenate(cols)
            _, comp = connected_components(coo_matrix((np.ones(len(r), np.int8), (r, c)), shape=(self.n, self.n)), directed=False)
            cores = np.flatnonzero(self.core)
            first = {}
            for i in cores:
                first[comp[i]] = min(first.get(comp[i], np.iinfo(np.int64).max), g[i])
            key[cores] = [first[comp[i]] for i in cores]
        return torch.from_numpy(key)

    def assign(self, core_label):
        cl = core_label.numpy()
        out = np.full(self.n, -1, dtype=np.int32)
        for i in range(self.n):
            if self.core[i]:
                out[i] = cl[i]
            else:
                nb = self.nbrs[i]
                nb = nb[self.core[nb]]
                if len(nb):
                    out[i] = cl[nb].min()
        return torch.from_numpy(out)


class OracleEngine(TorchEngineBase):
    """Numeric stages from the oracle; packing and collectives inherited (plain torch over gloo)."""
    device = torch.device("cpu")

    def __init__(self, spec):
        self.spec = spec

    def spoke_to_points(self, echo, cos_tab, sin_tab, range_res, sweep_gain, thr, stride, gpf, cap=None):
        W = echo.shape[0]
        xs, ys, zs, gs, off = [], [], [], [], [0]
        for w in range(W):
            x, y, z = O.sweep_to_points(echo[w].numpy(), self.spec.angle_units(), self.spec.scale(), thr, stride)
            xs.append(x); ys.append(y); zs.append(z); gs.append(np.full(len(x), int(sweep_gain[w]), np.int32))
            if (w + 1) % gpf == 0:
                off.append(off[-1] + sum(len(a) for a in xs[-gpf:]))
        cat = lambda parts, dt: torch.from_numpy(np.concatenate(parts).astype(dt))
        return PointBatch(cat(xs, np.float32), cat(ys, np.float32), cat(zs, np.float32), cat(gs, np.int32),
                          torch.tensor(off, dtype=torch.int64), off[-1])

    def bounds(self, x, y):
        return torch.tensor([x.min(), x.max(), y.min(), y.max()], dtype=torch.float32)

    def land_accumulate(self, batch, xe, ye):
        nx, ny = len(xe) - 1, len(ye) - 1
        count = np.zeros((nx, ny), np.int32)
        isum = np.zeros((nx, ny), np.float64)
        n = batch.n
        if n:
            ix = O.cell_index(batch.x[:n].numpy(), xe, nx)
            iy = O.cell_index(batch.y[:n].numpy(), ye, ny)
            np.add.at(count, (ix, iy), 1)
            np.add.at(isum, (ix, iy), batch.inten[:n].numpy())
        return torch.from_numpy(count), torch.from_numpy(isum)

    def land_cells(self, count, isum, built, persistence, min_intensity):
        return torch.from_numpy(O.land_cells(count.numpy(), isum.numpy(), built, persistence, min_intensity).astype(np.uint8))

    def land_filter(self, batch, xe, ye, land):
        n = batch.n
        pts = np.column_stack([batch.x[:n].numpy(), batch.y[:n].numpy()])
        keep = O.land_keep_mask(pts, land.numpy().astype(bool), (xe, ye))
        off = batch.frame_off.numpy()
        new_off = np.concatenate([[0], np.cumsum([keep[off[f]:off[f + 1]].sum() for f in range(len(off) - 1)])])
        k = torch.from_numpy(keep)
        return PointBatch(batch.x[:n][k], batch.y[:n][k], batch.inten[:n][k], batch.gain[:n][k],
                          torch.from_numpy(new_off.astype(np.int64)), -1)      # count read later, like the CUDA engine

    def expand_frame_times(self, frame_off, frame_ids, n):
        off = frame_off.numpy()
        return torch.from_numpy(np.repeat(frame_ids.numpy(), np.diff(off)).astype(np.float32))

    def phases(self, x, y, times, eps_space, eps_time, min_samples, hint=None):
        return OraclePhases(x, y, times, eps_space, eps_time, min_samples)

    def relabel(self, keys, table_keys, table_ids):
        k, tk, ti = keys.numpy(), table_keys.numpy(), table_ids.numpy()
        out = np.full(len(k), -1, np.int32)
        if len(tk):
            pos = np.clip(np.searchsorted(tk, k), 0, len(tk) - 1)
            hit = (k >= 0) & (tk[pos] == k)
            out[hit] = ti[pos[hit]]
        return torch.from_numpy(out)


# ------------------------------------------------------------------------------ single-process expectation
def expected(spec, cfg):
    echo = syn.synth_echo(spec)
    ang, scale = spec.angle_units(), spec.scale()
    frames = []
    for f in range(spec.frames):
        per_gain = {g: O.sweep_to_points(echo[f, gi], ang, scale, cfg.intensity_threshold, cfg.point_stride)
                    for gi, g in enumerate(spec.gains)}
        fused = O.fuse_concat(per_gain)
        frames.append(fused[0] if fused is not None else np.zeros((0, 3), np.float32))
    built = [p for p in frames if len(p)]
    if cfg.land_filter and len(built) > cfg.land_min_frames:
        count, isum, edges = O.occupancy_grid(built)
        land = O.land_cells(count, isum, len(built))
        frames = [p[O.land_keep_mask(p, land, edges)] if len(p) else p for p in frames]
    pts = np.concatenate(frames)
    fid = np.concatenate([np.full(len(p), i) for i, p in enumerate(frames)]).astype(np.float32)
    labels, _ = st_dbscan_c(pts[:, :2], fid, cfg.eps_space, cfg.eps_time, cfg.min_samples)
    return echo, frames, labels


def _worker(rank, world, port, spec_kw, cfg_kw, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = syn.SweepSpec(**spec_kw)
        cfg = DetectionConfig(**cfg_kw)
        echo = syn.synth_echo(spec)
        per = spec.frames // world
        lo, hi = rank * per, (spec.frames if rank == world - 1 else (rank + 1) * per)
        sd = ShardedDetection(cfg, engine=OracleEngine(spec))
        sd._key_cap = 3                      # far too small: exercises the collective "everybody repeats" path
        res = sd.run_device(torch.from_numpy(echo[lo:hi]), None, None, None, np.arange(lo, hi))
        n = res.points.n
        # the same block three times with two in flight (interleaved generators): identical results
        blk = (torch.from_numpy(echo[lo:hi]), None, None, None, np.arange(lo, hi))
        for again in sd.run_blocks([blk] * 3, in_flight=2):
            assert torch.equal(again.labels, res.labels) and again.n_clusters == res.n_clusters
        assert len(sd.run_blocks([blk] * 3, keep=False, in_flight=3)) == 1
        assert sd._key_cap > 3
        np.savez(Path(out_dir) / f"rank{rank}.npz", labels=res.labels.numpy(), x=res.points.x[:n].numpy(),
                 y=res.points.y[:n].numpy(), off=res.points.frame_off.numpy(), ncl=res.n_clusters, halo=np.array(res.halo_points))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


SPEC_LONG = dict(SPEC, frames=22, seed=23)       # 4 ranks x 5 frames (+2 on the last): room for a 5-frame halo


@pytest.mark.parametrize("world,cfg_kw,spec_kw", [(2, {}, SPEC), (3, {"eps_time": 1.0, "min_samples": 8}, SPEC),
                                                  (2, {"eps_time": 0.5, "land_filter": False}, SPEC),
                                                  (4, {"eps_time": 3.0, "min_samples": 10}, SPEC),      # shard = halo = 3 frames
                                                  (4, {"eps_time": 5.0}, SPEC_LONG)])                   # BASELINE config 5: halo 5 = shard
def test_sharded_equals_single_process(tmp_path, world, cfg_kw, spec_kw):
    cfg = DetectionConfig(**cfg_kw)
    spec = syn.SweepSpec(**spec_kw)
    _, frames, want = expected(spec, cfg)
    mp.spawn(_worker, args=(world, _free_port(), spec_kw, cfg_kw, str(tmp_path)), nprocs=world, join=True)
    got, xs, ncl = [], [], set()
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        got.append(d["labels"]); xs.append(d["x"]); ncl.add(int(d["ncl"]))
        if cfg.eps_time >= 1 and world > 1:
            assert d["halo"].sum() > 0                     # a halo was really exchanged
    want_x = np.concatenate([p[:, 0] for p in frames])
    assert np.array_equal(np.concatenate(xs), want_x)      # the land filter saw the global grid
    assert np.array_equal(np.concatenate(got), want)
    assert ncl == {int(want.max()) + 1}
    assert want.max() >= 2                                 # the case has clusters to stitch


def test_stitch_components_known_answer():
    e = np.zeros(0, np.int64)
    # rank 0 sees components keyed 5 and 40; rank 1 sees 38 (the same cluster as 40: rank 0's last-zone core point
    # keyed 40 is rank 1's left-halo point keyed 38) and 90
    segs = [[e, e, np.array([40]), np.array([40, 40])], [np.array([38]), np.array([38, 38]), e, e]]
    tk, ti, n = stitch_components(segs, [np.array([5, 40]), np.array([38, 90])])
    assert list(tk) == [5, 38, 40, 90] and list(ti) == [0, 1, 1, 2] and n == 3
    tk, ti, n = stitch_components([[e, e, e, e]], [e])
    assert len(tk) == 0 and n == 0


def _worker_sparse(rank, world, port, mode, out_dir):
    """A rank without any points (mode 'one_empty') / no points anywhere (mode 'all_empty')."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = syn.SweepSpec(**SPEC)
        cfg = DetectionConfig(land_filter=False)
        echo = syn.synth_echo(spec)
        per = spec.frames // world
        lo, hi = rank * per, (spec.frames if rank == world - 1 else (rank + 1) * per)
        mine = echo[lo:hi].copy()
        if mode == "all_empty" or rank == 1:
            mine[:] = 0.0                                   # nothing above the threshold on this rank
        sd = ShardedDetection(cfg, engine=OracleEngine(spec))
        res = sd.run_device(torch.from_numpy(mine), None, None, None, np.arange(lo, hi))
        many = sd.run_blocks([(torch.from_numpy(mine), None, None, None, np.arange(lo, hi))] * 2, in_flight=2)
        assert all(torch.equal(m.labels, res.labels) and m.n_clusters == res.n_clusters for m in many)
        np.savez(Path(out_dir) / f"rank{rank}.npz", labels=res.labels.numpy(), n=res.points.n, ncl=res.n_clusters)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["one_empty", "all_empty"])
def test_sharded_with_empty_ranks(tmp_path, mode):
    """Ranks without points must walk through the same collectives as the others (no hang, right labels)."""
    world = 3
    spec = syn.SweepSpec(**SPEC)
    cfg = DetectionConfig(land_filter=False)
    echo = syn.synth_echo(spec)
    per = spec.frames // world
    if mode == "all_empty":
        echo[:] = 0.0
    else:
        echo[per:2 * per] = 0.0                             # rank 1's frames
    ang, scale = spec.angle_units(), spec.scale()
    frames = []
    for f in range(spec.frames):
        per_gain = {g: O.sweep_to_points(echo[f, gi], ang, scale, cfg.intensity_threshold, cfg.point_stride) for gi, g in enumerate(spec.gains)}
        fused = O.fuse_concat(per_gain)
        frames.append(fused[0] if fused is not None else np.zeros((0, 3), np.float32))
    pts = np.concatenate(frames)
    fid = np.concatenate([np.full(len(p), i) for i, p in enumerate(frames)]).astype(np.float32)
    want = st_dbscan_c(pts[:, :2], fid, cfg.eps_space, cfg.eps_time, cfg.min_samples)[0] if len(pts) else np.zeros(0, np.int32)
    mp.spawn(_worker_sparse, args=(world, _free_port(), mode, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert np.array_equal(np.concatenate([g["labels"] for g in got]), want)
    assert {int(g["ncl"]) for g in got} == {int(want.max()) + 1 if len(want) and want.max() >= 0 else 0}
    if mode == "one_empty":
        assert int(got[1]["n"]) == 0 and want.max() >= 1


def test_run_blocks_schedule_is_deterministic_round_robin_with_skips():
    """The interleaving of blocks depends only on what the generators yield (constants), never on timing: that is what
    lets every rank issue its collectives in the same order. Scripted generators record the order of their steps."""
    spec = syn.SweepSpec(**SPEC)
    sd = ShardedDetection(DetectionConfig(), rank=0, world=1, engine=OracleEngine(spec))
    trace = []

    def scripted(name, yields):
        def gen(*_):
            for step, y in enumerate(yields):
                trace.append((name, step))
                yield None, y, []                                  # (event, rounds to sit out, host tensors)
            trace.append((name, "done"))
            return name
        return gen

    scripts = {"A": [3, 0, 0], "B": [0, 0], "C": [0]}
    order = iter(scripts)
    sd._run_gen = lambda *a: scripted(n := next(order), scripts[n])()
    out = sd.run_blocks([("A",), ("B",), ("C",)], in_flight=2, ordered=True)
    assert out == ["A", "B", "C"]                                   # results in block order
    # A steps once and sits out 3 rounds while B runs to its end; C takes B's slot; then A and C alternate
    assert trace == [("A", 0), ("B", 0), ("B", 1), ("B", "done"), ("C", 0), ("A", 1), ("C", "done"), ("A", 2), ("A", "done")]
    # a block alone never waits for its own skip count
    trace.clear()
    order = iter(["A"])
    assert sd.run_blocks([("A",)], in_flight=2, ordered=True) == ["A"]
    assert trace == [("A", 0), ("A", 1), ("A", 2), ("A", "done")]
    # unordered mode (the CUDA engine's default): a block is stepped whenever its read-back has arrived - with no events
    # pending that is plain round-robin, skip hints ignored; results still come back in block order
    trace.clear()
    order = iter(scripts)
    assert sd.run_blocks([("A",), ("B",), ("C",)], in_flight=2, ordered=False) == ["A", "B", "C"]
    assert trace == [("A", 0), ("B", 0), ("A", 1), ("B", 1), ("A", 2), ("B", "done"), ("A", "done"), ("C", 0), ("C", "done")]
