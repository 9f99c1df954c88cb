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
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import tracker as trk
from ._lib import RadarB200Error
from .device import PointBatch


# ------------------------------------------------------------------------------------------ stitching (host)
def stitch_components(seg_keys: Sequence[Sequence[np.ndarray]], keys: Sequence[np.ndarray]
                      ) -> Tuple[np.ndarray, np.ndarray, int]:
    """Global cluster numbering from the ranks' local components.

    ``keys[r]``: the distinct local component keys of rank r (key = smallest global core index the rank
    saw in that component). ``seg_keys[r]`` = four arrays with the local key of every CORE point of rank r's
    boundary zones, each in point order: ``[left halo, own first frames, own last frames, right halo]``.
    Rank r's "own last frames" and rank r+1's "left halo" are the same points with the same core flags, in
    the same order (likewise "right halo" / "own first frames"), so the two key arrays pair up position by
    position: each pair ties two local components together. Returns ``(table_keys sorted, table_ids,
    n_clusters)`` where the id of a component is the rank of its smallest global core index among all
    components — the reference's numbering."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    uniq = np.unique(np.concatenate([np.asarray(k, dtype=np.int64) for k in keys] + [np.zeros(0, np.int64)]))
    m = len(uniq)
    if m == 0:
        return uniq, np.zeros(0, np.int32), 0
    pa, pb = [np.zeros(0, np.int64)], [np.zeros(0, np.int64)]
    for r in range(len(seg_keys) - 1):
        for mine, theirs in ((seg_keys[r][2], seg_keys[r + 1][0]), (seg_keys[r][3], seg_keys[r + 1][1])):
            if len(mine) != len(theirs):
                raise RadarB200Error("stitch: boundary zones of neighbouring ranks do not match")
            pa.append(np.asarray(mine, dtype=np.int64)); pb.append(np.asarray(theirs, dtype=np.int64))
    a = np.searchsorted(uniq, np.concatenate(pa))
    b = np.searchsorted(uniq, np.concatenate(pb))
    pair = np.unique(a * m + b)                                 # few distinct component pairs
    a, b = pair // m, pair % m
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

    # -- the spoke stage of the NEXT block, launched on a side stream while the current block is in its exchange and
    #    clustering phases (it needs no exchange and no host knowledge): software pipelining inside one rank
    def spoke_launch(self, echo, cos_tab, sin_tab, range_res, sweep_gain, thr, stride, gpf, cap):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(self.device)
        main = torch.cuda.current_stream(self.device)
        # outputs are allocated on the MAIN stream (they are consumed and freed there; the caching allocator can then
        # recycle them block after block), only the launches go to the side stream
        W = echo.shape[0]
        out = (torch.empty(cap, dtype=torch.float32, device=self.device), torch.empty(cap, dtype=torch.float32, device=self.device),
               torch.empty(cap, dtype=torch.float32, device=self.device), torch.empty(cap, dtype=torch.int32, device=self.device),
               torch.empty(W + 1, dtype=torch.int64, device=self.device))
        self._side.wait_stream(main)                      # inputs (and the fresh buffers) are ready on the main stream
        with torch.cuda.stream(self._side):
            self.dev.spoke_to_points_raw(echo, cos_tab, sin_tab, range_res, sweep_gain, thr, stride, cap, out=out)
            done = torch.cuda.Event()
            done.record(self._side)
        return dict(outs=out, done=done, cap=cap, args=(echo, cos_tab, sin_tab, range_res, sweep_gain, thr, stride, gpf))

    def spoke_finish(self, h) -> PointBatch:
        torch.cuda.current_stream(self.device).wait_event(h["done"])
        x, y, inten, gain, sweep_base = h["outs"]
        n = int(sweep_base[-1].item())
        if n > h["cap"]:                                   # capacity guess too small: redo synchronously
            echo, cos_tab, sin_tab, range_res, sweep_gain, thr, stride, gpf = h["args"]
            return self.spoke_to_points(echo, cos_tab, sin_tab, range_res, sweep_gain, thr, stride, gpf, cap=n)
        gpf = h["args"][7]
        return PointBatch(x, y, inten, gain, self.dev.frame_offsets(sweep_base, (sweep_base.numel() - 1) // gpf, gpf), n)

    def bounds(self, x, y) -> torch.Tensor:
        return self.dev.bounds(x, y)

    def _edges(self, xe: np.ndarray, ye: np.ndarray):
        key = (xe.tobytes(), ye.tobytes())
        if getattr(self, "_edge_key", None) != key:              # one upload per distinct grid
            both = torch.from_numpy(np.concatenate([xe, ye])).to(self.device)
            self._edge_key, self._edge_dev = key, (both[:len(xe)], both[len(xe):])
        return self._edge_dev

    def land_accumulate(self, batch: PointBatch, xe: np.ndarray, ye: np.ndarray):
        d_xe, d_ye = self._edges(xe, ye)
        n = batch.n
        if n == 0:
            return (torch.zeros((len(xe) - 1, len(ye) - 1), dtype=torch.int32, device=self.device),
                    torch.zeros((len(xe) - 1, len(ye) - 1), dtype=torch.float64, device=self.device))
        return self.dev.land_accumulate(batch.x[:n], batch.y[:n], batch.inten[:n], d_xe, d_ye)

    def land_cells(self, count, isum, built, persistence, min_intensity):
        return self.dev.land_cells(count, isum, built, persistence, min_intensity)

    def land_filter(self, batch: PointBatch, xe, ye, land) -> PointBatch:
        d_xe, d_ye = self._edges(xe, ye)
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
        self._gain_cache = None
        self._prefetched = None
        self.profile = False               # True: synchronise and record wall-clock per stage in self.timings
        self.timings = {}
        self._t_last = None

    def _tick(self, name: Optional[str]) -> None:
        """Stage timer (diagnostics): time since the previous tick is booked under ``name``."""
        if not self.profile:
            return
        import time
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        now = time.perf_counter()
        if name is not None and self._t_last is not None:
            self.timings[name] = self.timings.get(name, 0.0) + (now - self._t_last)
        self._t_last = now

    # ---- small collective helpers (tensors live on the engine's device) -----------------------------
    def _t(self, values, dtype) -> torch.Tensor:
        return torch.tensor(values, dtype=dtype, device=self.device)

    def _all_gather_vec(self, vec: np.ndarray, dtype=torch.float64) -> np.ndarray:
        """All-gather one small fixed-length vector per rank -> ``[world, len]`` on the host (ONE collective)."""
        mine = torch.from_numpy(np.ascontiguousarray(vec)).to(dtype).to(self.device)
        if self.world == 1:
            return mine.cpu().numpy()[None]
        out = torch.empty((self.world, mine.numel()), dtype=dtype, device=self.device)
        dist.all_gather_into_tensor(out, mine[None], group=self.group)
        return out.cpu().numpy()

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
    def _gains(self, n_sweeps: int) -> torch.Tensor:
        if self._gain_cache is None or self._gain_cache.numel() != n_sweeps:
            self._gain_cache = self._t(list(self.cfg.gains) * (n_sweeps // len(self.cfg.gains)), torch.int32)
        return self._gain_cache

    def prefetch(self, echo, cos_tab, sin_tab, range_res) -> None:
        """Launch the spoke-to-point stage of the block that :meth:`run_device` will be called with NEXT, on a side
        stream, without waiting for it: it overlaps the exchange and clustering phases of the current block.
        Optional; engines without ``spoke_launch`` ignore it."""
        if not hasattr(self.engine, "spoke_launch"):
            return
        F, G, S, E = echo.shape
        cap = self._cap_hint or max(1, F * G * ((S * E + self.cfg.point_stride - 1) // max(self.cfg.point_stride, 1)) // 8)
        h = self.engine.spoke_launch(echo.reshape(F * G, S, E), cos_tab, sin_tab, range_res, self._gains(F * G),
                                     self.cfg.intensity_threshold, self.cfg.point_stride, G, cap)
        self._prefetched = (echo.data_ptr(), tuple(echo.shape), h)

    def run_blocks(self, blocks, keep: bool = True):
        """Run a sequence of blocks ``(echo, cos_tab, sin_tab, range_res, frame_ids)`` of this rank, software
        pipelined: while block k goes through its exchange and clustering phases, the spoke stage of block k+1
        is already running on the side stream. Same results as calling :meth:`run_device` per block.
        Returns the results (``keep=False``: only the last one, earlier ones are released for buffer reuse)."""
        blocks = list(blocks)
        out, nxt = [], None
        for i, blk in enumerate(blocks):
            if nxt is None:
                self.prefetch(*blk[:4])
                nxt = self._prefetched
            cur, self._prefetched = nxt, None
            nxt = None
            if i + 1 < len(blocks):
                self.prefetch(*blocks[i + 1][:4])              # goes out BEFORE block i's phases are enqueued
                nxt = self._prefetched
            self._prefetched = cur
            res = self.run_device(*blk)
            if keep or i == len(blocks) - 1:
                out.append(res)
            del res
        return out

    def run_device(self, echo, cos_tab, sin_tab, range_res, frame_ids: Sequence[int], cluster: bool = True) -> ShardResult:
        """``echo[F,G,S,E]`` = this rank's frames (device tensor of the engine); ``frame_ids`` their ids,
        increasing across ranks (rank r's ids are all smaller than rank r+1's)."""
        cfg, eng = self.cfg, self.engine
        F, G, S, E = echo.shape
        ids = np.asarray(frame_ids, dtype=np.int64)
        self._tick(None)
        pf, self._prefetched = self._prefetched, None
        if pf is not None and pf[0] == echo.data_ptr() and pf[1] == tuple(echo.shape):
            raw = eng.spoke_finish(pf[2])
        else:
            raw = eng.spoke_to_points(echo.reshape(F * G, S, E), cos_tab, sin_tab, range_res, self._gains(F * G),
                                      cfg.intensity_threshold, cfg.point_stride, G, cap=self._cap_hint or None)
        self._cap_hint = max(self._cap_hint, int(raw.n * 1.25) + 1024)
        self._tick("spoke")

        # ---- land / stationary persistence filter over ALL ranks' frames --------------------------
        # collective 1: [frames built, points, xmin, xmax, ymin, ymax] of every rank (float64 holds them exactly)
        pts, land, edges = raw, None, None
        if cfg.land_filter:
            raw_off = raw.frame_off.cpu().numpy()
            mine = np.zeros(6, dtype=np.float64)
            mine[0], mine[1] = np.count_nonzero(np.diff(raw_off)), raw.n
            if raw.n > 0:
                mine[2:] = eng.bounds(raw.x[:raw.n], raw.y[:raw.n]).cpu().numpy().astype(np.float64)
            allv = self._all_gather_vec(mine)
            built, n_all = int(allv[:, 0].sum()), int(allv[:, 1].sum())
            if n_all > 0 and built > cfg.land_min_frames:
                have = allv[allv[:, 1] > 0]
                b4 = np.array([have[:, 2].min(), have[:, 3].max(), have[:, 4].min(), have[:, 5].max()], dtype=np.float32)
                xe, ye = trk.grid_edges_from_bounds(b4, cfg.land_resolution)
                count, isum = eng.land_accumulate(raw, xe, ye)
                if self.world > 1:
                    # collective 2: counts and sums in ONE float64 all-reduce (counts < 2^53 stay exact)
                    grids = torch.stack([count.to(torch.float64), isum])
                    dist.all_reduce(grids, op=dist.ReduceOp.SUM, group=self.group)
                    count, isum = grids[0].to(torch.int32), grids[1].contiguous()
                land = eng.land_cells(count, isum, built, cfg.land_persistence, cfg.land_min_intensity)
                if raw.n > 0:
                    pts = eng.land_filter(raw, xe, ye, land)
                edges = (xe, ye)

        self._tick("land")
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
        if self.world > 1 and h > F:
            raise RadarB200Error(f"time shard of {F} frames is shorter than the eps_time halo ({h} frames)")
        hh = min(h, F)
        off = pts.frame_off.cpu().numpy().astype(np.int64)
        lo_end, hi_start = int(off[hh]), int(off[F - hh])          # owned points [0, lo_end) go left, [hi_start, n) right

        # collective 3: per rank [points owned, start of the last-hh-frames zone, ids and per-frame counts of the
        # first hh and last hh frames] -> global index bases and both neighbours' halo layouts at once
        mine = np.concatenate([[n_own, hi_start], ids[:hh], np.diff(off)[:hh], ids[F - hh:], np.diff(off)[F - hh:]]).astype(np.int64)
        meta = self._all_gather_vec(mine, torch.int64)
        bases = np.concatenate([[0], np.cumsum(meta[:, 0])])
        gbase = int(bases[self.rank])
        lids = lcnt = rids = rcnt = np.zeros(0, np.int64)
        lbase = rbase = 0
        if self.rank > 0 and hh:                                    # left halo = the left neighbour's LAST hh frames
            m = meta[self.rank - 1]
            lids, lcnt, lbase = m[2 + 2 * hh:2 + 3 * hh], m[2 + 3 * hh:2 + 4 * hh], int(bases[self.rank - 1] + m[1])
        if self.rank < self.world - 1 and hh:                       # right halo = the right neighbour's FIRST hh frames
            m = meta[self.rank + 1]
            rids, rcnt, rbase = m[2:2 + hh], m[2 + hh:2 + 2 * hh], int(bases[self.rank + 1])
        nl, nr = int(lcnt.sum()), int(rcnt.sum())
        n_loc = nl + n_own + nr
        if int(meta[:, 0].sum()) == 0:
            return torch.empty(0, dtype=torch.int32, device=self.device), 0, (nl, nr)

        xy = torch.stack([pts.x[:n_own], pts.y[:n_own]]) if n_own else torch.zeros((2, 0), dtype=torch.float32, device=self.device)
        xl, xr = self._exchange(xy[:, :lo_end], xy[:, hi_start:], nl, nr, torch.float32, width=2)
        self._tick("halo")

        # ---- local problem: [left halo | owned | right halo], times = frame ids --------------------------
        XY = torch.cat([t for t in (xl, xy, xr) if t is not None], dim=1)
        X, Y = XY[0].contiguous(), XY[1].contiguous()
        all_ids = np.concatenate([lids, ids, rids]).astype(np.float32)
        all_cnt = np.concatenate([lcnt, np.diff(off), rcnt]).astype(np.int64)
        head = np.concatenate([[0], np.cumsum(all_cnt)]).astype(np.int64)
        aux = torch.from_numpy(np.concatenate([head, all_ids.astype(np.int64)])).to(self.device)      # one upload
        loc_off = aux[:len(head)]
        times = eng.expand_frame_times(loc_off, aux[len(head):].to(torch.float32), n_loc) if n_loc else \
            torch.zeros(0, dtype=torch.float32, device=self.device)
        gidx = torch.arange(n_loc, dtype=torch.int64, device=self.device)
        gidx[:nl] += lbase
        gidx[nl:nl + n_own] += gbase - nl
        gidx[nl + n_own:] += rbase - nl - n_own
        self._tick("prep")

        ph = eng.phases(X, Y, times, cfg.eps_space, cfg.eps_time, cfg.min_samples) if n_loc else None
        core = ph.cores() if n_loc else torch.zeros(0, dtype=torch.uint8, device=self.device)
        self._tick("plan+cores")
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
        self._tick("core-exchange+components")

        # ---- stitch on rank 0 ---------------------------------------------------------------------------------
        # what rank 0 needs from me: the local key of every CORE point of my four boundary zones (left halo, own
        # first frames, own last frames, right halo - they pair up with the neighbours' zones by position) and my
        # distinct component keys (= keys of the points that ARE their component's smallest core).
        # One device->host read, then collectives 4 (sizes) and 5 (packed gather).
        if n_loc:
            zones = ((0, nl), (nl, nl + lo_end), (nl + hi_start, nl + n_own), (nl + n_own, n_loc))
            is_core = key >= 0
            seg_counts = [int(c) for c in torch.stack([is_core[a:b].sum() for a, b in zones]).cpu().numpy()]
            parts = [key[a:b][is_core[a:b]] for a, b in zones] + [key[key == gidx]]
            flat = torch.cat(parts).cpu().numpy()
            segs = np.split(flat, np.cumsum(seg_counts))
            my_segs, my_keys = segs[:4], segs[4]
        else:
            my_segs, my_keys = [np.zeros(0, np.int64)] * 4, np.zeros(0, np.int64)
        self._tick("select")
        packed = np.concatenate(list(my_segs) + [my_keys]).astype(np.int64)
        sizes = self._all_gather_vec(np.array([len(x) for x in my_segs] + [len(my_keys)], dtype=np.int64), torch.int64)
        cap_keys = int(sizes[:, 4].sum())
        table = np.zeros(2 + 2 * cap_keys, dtype=np.int64)
        if self.world == 1:
            tk, ti, ncl = stitch_components([my_segs], [my_keys])
        else:
            cap = int(sizes.sum(axis=1).max())
            pad = torch.zeros(max(cap, 1), dtype=torch.int64, device=self.device)
            pad[:len(packed)] = torch.from_numpy(packed).to(self.device)
            bufs = [torch.empty_like(pad) for _ in range(self.world)] if self.rank == 0 else None
            dist.gather(pad, bufs, dst=0, group=self.group)
            self._tick("gather")
            if self.rank == 0:
                got = torch.stack(bufs).cpu().numpy()
                all_segs, all_keys = [], []
                for r in range(self.world):
                    pieces = np.split(got[r, :int(sizes[r].sum())], np.cumsum(sizes[r])[:-1])
                    all_segs.append(pieces[:4]); all_keys.append(pieces[4])
                tk, ti, ncl = stitch_components(all_segs, all_keys)
                table[0], table[1] = len(tk), ncl
                table[2:2 + len(tk)] = tk
                table[2 + cap_keys:2 + cap_keys + len(tk)] = ti
            self._tick("stitch")
            # collective 6: the key -> id table
            t_table = torch.from_numpy(table).to(self.device)
            dist.broadcast(t_table, src=0, group=self.group)
            table = t_table.cpu().numpy()
            m, ncl = int(table[0]), int(table[1])
            tk, ti = table[2:2 + m], table[2 + cap_keys:2 + cap_keys + m].astype(np.int32)
            self._tick("broadcast")

        # ---- labels of the owned points -------------------------------------------------------------------------
        if n_loc == 0:
            return torch.empty(0, dtype=torch.int32, device=self.device), int(ncl), (nl, nr)
        core_label = eng.relabel(key, torch.from_numpy(np.ascontiguousarray(tk)).to(self.device),
                                 torch.from_numpy(np.ascontiguousarray(ti, dtype=np.int32)).to(self.device))
        labels = ph.assign(core_label)
        out = labels[nl:nl + n_own].contiguous()
        self._tick("relabel+assign")
        return out, int(ncl), (nl, nr)

    # ---- host entry (mirrors DetectionPipeline.run_host for this rank's block) -------------------------------
    def run_host(self, echo, angle_units, scale, frame_ids: Sequence[int], pinned: Optional[torch.Tensor] = None) -> dict:
        """Host buffers in, host buffers out for this rank's frames: uploads ``echo[F,G,S,E]`` (numpy or a
        pinned torch tensor), runs the sharded device path and reads this rank's result back."""
        if self.base is None:
            raise RadarB200Error("run_host needs the CUDA engine")
        if pinned is not None:
            t_echo = pinned
        else:
            a = np.asarray(echo)
            t_echo = torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint8 if a.dtype == np.uint8 else np.float32))
        F, G, S, E = t_echo.shape
        c, s, r = self.base.spoke_tables(angle_units, scale, F, E)
        d = self.device
        d_echo = t_echo.to(d, non_blocking=True)
        res = self.run_device(d_echo, torch.from_numpy(c).to(d), torch.from_numpy(s).to(d), torch.from_numpy(r).to(d), frame_ids)
        out = res.to_host()
        out["n_clusters"] = res.n_clusters
        out["h2d_bytes"] = t_echo.numel() * t_echo.element_size() + 3 * c.nbytes
        out["d2h_bytes"] = out["points"].nbytes + out["gains"].nbytes + out["labels"].nbytes + out["frame_off"].nbytes
        return out

