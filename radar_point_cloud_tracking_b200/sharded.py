"""Time-sharded multi-GPU driver of the detection path (SURVEY.md section 8 e): one process per GPU,
``torch.distributed`` (NCCL over NVLink/NVSwitch; gloo in the CPU tests) for the few small exchanges.

Rank r owns a contiguous block of frames. Spoke-to-point is embarrassingly parallel; the land filter
needs two tiny reductions; ST-DBSCAN couples frame f only to frames within ``eps_time``, so every rank
clusters its own frames plus a ``floor(eps_time)``-frame halo from each neighbour and the cluster ids are
stitched on rank 0:

1. ``all_reduce(MIN)`` of ``[xmin, -xmax, ymin, -ymax]``            -> identical ``np.arange`` edges everywhere
2. ``all_reduce(SUM)`` of the per-cell point counts / intensity sums -> identical land mask everywhere
3. neighbour send/recv of the boundary frames' filtered points (x, y) -> exact core flags of OWNED points
4. neighbour send/recv of those points' core flags (1 B/point)        -> exact core flags of HALO points
5. components over owned + halo cores, keyed by GLOBAL point index; gather of
   (global index, local component key) of every boundary-zone core point to rank 0, union of the keys
   that share a point, canonical numbering (rank of the smallest core index, exactly the reference's),
   ``broadcast`` of the key -> id table
6. ``rb_relabel`` + border assignment (all neighbours of an owned point are present locally)

The result equals the single-GPU labels of the concatenated recording, id for id.

The collectives and the stitching are host logic and device-agnostic torch plumbing; every numeric
stage goes through an *engine* (:class:`CudaEngine` = the C ABI). The CPU tests drive the same logic
over gloo with an oracle-backed engine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import tracker as trk
from ._lib import RadarB200Error
from .device import PointBatch


# ------------------------------------------------------------------------------------------ stitching (host)
def stitch_components(link_gidx: np.ndarray, link_key: np.ndarray, keys: np.ndarray
                      ) -> Tuple[np.ndarray, np.ndarray, int]:
    """Global cluster numbering from the ranks' local components.

    ``keys``: every local component key of every rank (key = smallest global core index the rank saw
    in that component). ``link_*``: for every copy of a boundary-zone core point (its owner's copy and
    the neighbour's halo copy) the point's global index and the key of the local component it is in.
    Copies of one point tie their components together. Returns ``(table_keys sorted, table_ids,
    n_clusters)`` where the id of a component is the rank of its smallest global core index among
    all components — the reference's numbering."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    link_gidx = np.asarray(link_gidx, dtype=np.int64)
    link_key = np.asarray(link_key, dtype=np.int64)
    uniq = np.unique(np.concatenate([np.asarray(keys, dtype=np.int64), link_key]))
    m = len(uniq)
    if m == 0:
        return uniq, np.zeros(0, np.int32), 0
    node = np.searchsorted(uniq, link_key)
    order = np.argsort(link_gidx, kind="stable")
    g, nd = link_gidx[order], node[order]
    same = g[1:] == g[:-1]
    a, b = nd[:-1][same], nd[1:][same]
    graph = coo_matrix((np.ones(len(a), np.int8), (a, b)), shape=(m, m))
    ncomp, comp = connected_components(graph, directed=False)
    first = np.full(ncomp, m, dtype=np.int64)
    np.minimum.at(first, comp, np.arange(m))                  # uniq is sorted: smallest node = smallest key
    final_key = uniq[first[comp]]
    finals = np.unique(final_key)
    ids = np.searchsorted(finals, final_key).astype(np.int32)
    return uniq, ids, len(finals)


# ------------------------------------------------------------------------------------------ engines
class CudaEngine:
    """Numeric stages on the current CUDA device through ``libradarb200.so``."""

    def __init__(self, device_index: Optional[int] = None):
        from . import device as dev

        if not torch.cuda.is_available():
            raise RadarB200Error("no CUDA device: the sharded detection path is GPU only (no CPU fallback)")
        self.dev = dev
        self.device = torch.device("cuda", torch.cuda.current_device() if device_index is None else device_index)

    def spoke_to_points(self, echo, cos_tab, sin_tab, range_res, sweep_gain, thr, stride, gpf, cap=None) -> PointBatch:
        return self.dev.spoke_to_points(echo, cos_tab, sin_tab, range_res, sweep_gain, thr, stride, gains_per_frame=gpf, cap=cap)

    def bounds(self, x, y) -> torch.Tensor:
        return self.dev.bounds(x, y)

    def land_accumulate(self, batch: PointBatch, xe: np.ndarray, ye: np.ndarray):
        d_xe, d_ye = torch.from_numpy(xe).to(self.device), torch.from_numpy(ye).to(self.device)
        n = batch.n
        if n == 0:
            return (torch.zeros((len(xe) - 1, len(ye) - 1), dtype=torch.int32, device=self.device),
                    torch.zeros((len(xe) - 1, len(ye) - 1), dtype=torch.float64, device=self.device))
        return self.dev.land_accumulate(batch.x[:n], batch.y[:n], batch.inten[:n], d_xe, d_ye)

    def land_cells(self, count, isum, built, persistence, min_intensity):
        return self.dev.land_cells(count, isum, built, persistence, min_intensity)

    def land_filter(self, batch: PointBatch, xe, ye, land) -> PointBatch:
        d_xe, d_ye = torch.from_numpy(xe).to(self.device), torch.from_numpy(ye).to(self.device)
        return self.dev.land_filter(batch, d_xe, d_ye, land)

    def expand_frame_times(self, frame_off, frame_ids, n):
        return self.dev.expand_frame_times(frame_off, frame_ids, n)

    def phases(self, x, y, times, eps_space, eps_time, min_samples):
        return self.dev.StDbscanPhases(x, y, None, times, eps_space, eps_time, min_samples, stride=1, n=times.numel())

    def relabel(self, keys, table_keys, table_ids):
        return self.dev.relabel(keys, table_keys, table_ids)


# ------------------------------------------------------------------------------------------ result
@dataclass
class ShardResult:
    """Result for the frames owned by this rank (same fields the single-GPU ``DetectionResult`` has)."""
    frame_ids: np.ndarray
    raw: PointBatch
    points: PointBatch
    labels: torch.Tensor
    n_clusters: int                        # GLOBAL number of clusters
    halo_points: Tuple[int, int] = (0, 0)  # points received from the left / right neighbour
    land: Optional[torch.Tensor] = None
    edges: Optional[Tuple[np.ndarray, np.ndarray]] = None

    def to_host(self) -> dict:
        n, p = self.points.n, self.points
        packed = torch.stack([p.x[:n], p.y[:n], p.inten[:n]], dim=1)
        return dict(points=packed.cpu().numpy(), gains=p.gain[:n].cpu().numpy(), frame_off=p.frame_off.cpu().numpy(),
                    labels=self.labels[:n].cpu().numpy(), frame_ids=self.frame_ids)


# ------------------------------------------------------------------------------------------ driver
class ShardedDetection:
    """Runs the hot path for this rank's block of frames, exchanging what the stages need with the
    other ranks. ``engine`` defaults to :class:`CudaEngine`; ``group`` to the default process group."""

    def __init__(self, config=None, rank: Optional[int] = None, world: Optional[int] = None,
                 device: Optional[int] = None, engine=None, group=None):
        from .pipeline import DetectionConfig

        self.cfg = config or DetectionConfig()
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.engine = engine or CudaEngine(device)
        self.device = self.engine.device
        self.base = None
        if engine is None:
            from .pipeline import DetectionPipeline
            self.base = DetectionPipeline(self.cfg, self.device.index)       # spoke tables + ctx for the bench
        self._cap_hint = 0

    # ---- small collective helpers (tensors live on the engine's device) -----------------------------
    def _t(self, values, dtype) -> torch.Tensor:
        return torch.tensor(values, dtype=dtype, device=self.device)

    def _all_reduce(self, t: torch.Tensor, op) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=op, group=self.group)
        return t

    def _all_gather_i64(self, value: int) -> np.ndarray:
        if self.world == 1:
            return np.array([value], dtype=np.int64)
        out = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, self._t([value], torch.int64), group=self.group)
        return out.cpu().numpy()

    def _gather_varlen(self, t: torch.Tensor) -> Optional[List[np.ndarray]]:
        """Gather 1-D int64 tensors of different lengths to rank 0 (returns the list there, else None)."""
        sizes = self._all_gather_i64(t.numel())
        if self.world == 1:
            return [t.cpu().numpy()]
        cap = int(sizes.max())
        pad = torch.zeros(max(cap, 1), dtype=torch.int64, device=self.device)
        pad[:t.numel()] = t
        bufs = [torch.zeros_like(pad) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(pad, bufs, dst=0, group=self.group)
        if self.rank != 0:
            return None
        return [b[:int(s)].cpu().numpy() for b, s in zip(bufs, sizes)]

    def _exchange(self, to_left: Optional[torch.Tensor], to_right: Optional[torch.Tensor],
                  n_from_left: int, n_from_right: int, dtype, width: int = 1):
        """Send 2-D ``[width, n]`` tensors to the neighbours and receive theirs (sizes known)."""
        ops, left, right = [], None, None
        if self.rank > 0:
            if to_left is not None and to_left.numel():
                ops.append(dist.P2POp(dist.isend, to_left.contiguous(), self.rank - 1, group=self.group))
            if n_from_left:
                left = torch.empty((width, n_from_left), dtype=dtype, device=self.device)
                ops.append(dist.P2POp(dist.irecv, left, self.rank - 1, group=self.group))
        if self.rank < self.world - 1:
            if to_right is not None and to_right.numel():
                ops.append(dist.P2POp(dist.isend, to_right.contiguous(), self.rank + 1, group=self.group))
            if n_from_right:
                right = torch.empty((width, n_from_right), dtype=dtype, device=self.device)
                ops.append(dist.P2POp(dist.irecv, right, self.rank + 1, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return left, right

    # ---- the path -------------------------------------------------------------------------------------
    def run_device(self, echo, cos_tab, sin_tab, range_res, frame_ids: Sequence[int], cluster: bool = True) -> ShardResult:
        """``echo[F,G,S,E]`` = this rank's frames (device tensor of the engine); ``frame_ids`` their ids,
        increasing across ranks (rank r's ids are all smaller than rank r+1's)."""
        cfg, eng = self.cfg, self.engine
        F, G, S, E = echo.shape
        ids = np.asarray(frame_ids, dtype=np.int64)
        sweep_gain = self._t(list(cfg.gains) * F, torch.int32)
        raw = eng.spoke_to_points(echo.reshape(F * G, S, E), cos_tab, sin_tab, range_res, sweep_gain,
                                  cfg.intensity_threshold, cfg.point_stride, G, cap=self._cap_hint or None)
        self._cap_hint = max(self._cap_hint, int(raw.n * 1.25) + 1024)

        # ---- land / stationary persistence filter over ALL ranks' frames --------------------------
        pts, land, edges = raw, None, None
        if cfg.land_filter:
            raw_off = raw.frame_off.cpu().numpy()
            stats = self._all_reduce(self._t([int(np.count_nonzero(np.diff(raw_off))), raw.n], torch.int64), dist.ReduceOp.SUM)
            built, n_all = (int(v) for v in stats.cpu().numpy())
            if n_all > 0 and built > cfg.land_min_frames:
                if raw.n > 0:
                    b = eng.bounds(raw.x[:raw.n], raw.y[:raw.n]).to(torch.float32)
                    packed = torch.stack([b[0], -b[1], b[2], -b[3]])
                else:
                    packed = torch.full((4,), float("inf"), dtype=torch.float32, device=self.device)
                packed = self._all_reduce(packed, dist.ReduceOp.MIN).cpu().numpy()
                b4 = np.array([packed[0], -packed[1], packed[2], -packed[3]], dtype=np.float32)
                xe, ye = trk.grid_edges_from_bounds(b4, cfg.land_resolution)
                count, isum = eng.land_accumulate(raw, xe, ye)
                self._all_reduce(count, dist.ReduceOp.SUM)
                self._all_reduce(isum, dist.ReduceOp.SUM)
                land = eng.land_cells(count, isum, built, cfg.land_persistence, cfg.land_min_intensity)
                if raw.n > 0:
                    pts = eng.land_filter(raw, xe, ye, land)
                edges = (xe, ye)

        labels = torch.empty(0, dtype=torch.int32, device=self.device)
        n_clusters, halo = 0, (0, 0)
        if cluster:
            labels, n_clusters, halo = self._cluster(pts, ids)
        return ShardResult(ids, raw, pts, labels, n_clusters, halo, land, edges)

    def _cluster(self, pts: PointBatch, ids: np.ndarray):
        cfg, eng = self.cfg, self.engine
        F = len(ids)
        n_own = pts.n
        h = int(math.floor(cfg.eps_time)) if cfg.eps_time >= 0 else 0
        hh = min(h, F)
        if self.world > 1 and h > F:
            raise RadarB200Error(f"time shard of {F} frames is shorter than the eps_time halo ({h} frames)")
        off = pts.frame_off.cpu().numpy().astype(np.int64)
        counts = self._all_gather_i64(n_own)
        gbase = int(counts[:self.rank].sum())                      # global index of my first owned point
        lo_end, hi_start = int(off[hh]), int(off[F - hh])          # owned points [0, lo_end) go left, [hi_start, n) right

        # ---- halo: headers (frame ids + per-frame counts + global base), then x/y ---------------------
        def header(first_frame: int, start: int) -> torch.Tensor:
            fr = np.arange(first_frame, first_frame + hh)
            return self._t([gbase + start] + [int(ids[f]) for f in fr] + [int(off[f + 1] - off[f]) for f in fr], torch.int64)[None]

        hl, hr = self._exchange(header(0, 0) if hh else None, header(F - hh, hi_start) if hh else None,
                                (1 + 2 * hh) if hh else 0, (1 + 2 * hh) if hh else 0, torch.int64)
        def parse(hd):
            if hd is None:
                return 0, np.zeros(0, np.int64), np.zeros(0, np.int64)
            v = hd[0].cpu().numpy()
            return int(v[0]), v[1:1 + hh], v[1 + hh:1 + 2 * hh]
        lbase, lids, lcnt = parse(hl)
        rbase, rids, rcnt = parse(hr)
        nl, nr = int(lcnt.sum()), int(rcnt.sum())
        xy = torch.stack([pts.x[:n_own], pts.y[:n_own]]) if n_own else torch.zeros((2, 0), dtype=torch.float32, device=self.device)
        xl, xr = self._exchange(xy[:, :lo_end], xy[:, hi_start:], nl, nr, torch.float32, width=2)

        # ---- local problem: [left halo | owned | right halo], times = frame ids --------------------------
        parts_x = [t[0] for t in (xl,) if t is not None] + [xy[0]] + [t[0] for t in (xr,) if t is not None]
        parts_y = [t[1] for t in (xl,) if t is not None] + [xy[1]] + [t[1] for t in (xr,) if t is not None]
        X, Y = torch.cat(parts_x), torch.cat(parts_y)
        n_loc = nl + n_own + nr
        if n_loc == 0 or self._all_gather_i64(n_loc).sum() == 0:
            return torch.empty(0, dtype=torch.int32, device=self.device), 0, (nl, nr)
        all_ids = np.concatenate([lids, ids, rids]).astype(np.float32)
        all_cnt = np.concatenate([lcnt, np.diff(off), rcnt]).astype(np.int64)
        loc_off = self._t(np.concatenate([[0], np.cumsum(all_cnt)]).tolist(), torch.int64)
        times = eng.expand_frame_times(loc_off, torch.from_numpy(all_ids).to(self.device), n_loc)
        gidx = torch.cat([torch.arange(lbase, lbase + nl, dtype=torch.int64, device=self.device),
                          torch.arange(gbase, gbase + n_own, dtype=torch.int64, device=self.device),
                          torch.arange(rbase, rbase + nr, dtype=torch.int64, device=self.device)])

        ph = eng.phases(X, Y, times, cfg.eps_space, cfg.eps_time, cfg.min_samples) if n_loc else None
        core = ph.cores() if n_loc else torch.zeros(0, dtype=torch.uint8, device=self.device)
        # ---- exact core flags of the halo points come from their owners ----------------------------------
        own_core = core[nl:nl + n_own]
        cl, cr = self._exchange(own_core[None, :lo_end], own_core[None, hi_start:], nl, nr, torch.uint8)
        if n_loc:
            if cl is not None:
                core[:nl] = cl[0]
            if cr is not None:
                core[nl + n_own:] = cr[0]
            ph.set_cores(core)
            key = ph.components(gidx)
        else:
            key = torch.zeros(0, dtype=torch.int64, device=self.device)

        # ---- stitch on rank 0 ---------------------------------------------------------------------------------
        zone = torch.zeros(n_loc, dtype=torch.bool, device=self.device)
        zone[:nl + lo_end] = True                                   # left halo + my first frames
        zone[nl + hi_start:] = True                                 # my last frames + right halo
        link = zone & (key >= 0)
        comp_keys = torch.unique(key[key >= 0]) if n_loc else key
        g_gidx = self._gather_varlen(gidx[link])
        g_key = self._gather_varlen(key[link])
        g_all = self._gather_varlen(comp_keys)
        if self.rank == 0:
            tk, ti, ncl = stitch_components(np.concatenate(g_gidx), np.concatenate(g_key), np.concatenate(g_all))
            meta = self._t([len(tk), ncl], torch.int64)
        else:
            meta = torch.zeros(2, dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.broadcast(meta, src=0, group=self.group)
        m, ncl = (int(v) for v in meta.cpu().numpy())
        if self.rank == 0:
            table_k = torch.from_numpy(tk).to(self.device)
            table_i = torch.from_numpy(ti).to(self.device)
        else:
            table_k = torch.zeros(m, dtype=torch.int64, device=self.device)
            table_i = torch.zeros(m, dtype=torch.int32, device=self.device)
        if self.world > 1 and m > 0:
            dist.broadcast(table_k, src=0, group=self.group)
            dist.broadcast(table_i, src=0, group=self.group)

        # ---- labels of the owned points -------------------------------------------------------------------------
        if n_loc == 0:
            return torch.empty(0, dtype=torch.int32, device=self.device), ncl, (nl, nr)
        core_label = eng.relabel(key, table_k, table_i)
        labels = ph.assign(core_label)
        return labels[nl:nl + n_own].contiguous(), ncl, (nl, nr)

    # ---- host entry (mirrors DetectionPipeline.run_host for this rank's block) -------------------------------
    def run_host(self, echo, angle_units, scale, frame_ids: Sequence[int], pinned: Optional[torch.Tensor] = None) -> dict:
        """Host buffers in, host buffers out for this rank's frames: uploads ``echo[F,G,S,E]`` (numpy or a
        pinned torch tensor), runs the sharded device path and reads this rank's result back."""
        if self.base is None:
            raise RadarB200Error("run_host needs the CUDA engine")
        t_echo = pinned if pinned is not None else torch.from_numpy(np.ascontiguousarray(echo, dtype=np.float32))
        F, G, S, E = t_echo.shape
        c, s, r = self.base.spoke_tables(angle_units, scale, F, E)
        d = self.device
        d_echo = t_echo.to(d, non_blocking=True)
        res = self.run_device(d_echo, torch.from_numpy(c).to(d), torch.from_numpy(s).to(d), torch.from_numpy(r).to(d), frame_ids)
        out = res.to_host()
        out["n_clusters"] = res.n_clusters
        out["h2d_bytes"] = t_echo.numel() * 4 + 3 * c.nbytes
        out["d2h_bytes"] = out["points"].nbytes + out["gains"].nbytes + out["labels"].nbytes + out["frame_off"].nbytes
        return out
