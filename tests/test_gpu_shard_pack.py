"""The device-side bookkeeping of the time-sharded driver (csrc/shard.cu: rb_shard_pack_stats / _pack_layout /
_local_index / _pack_keys) against the plain-torch statement of the same vectors (sharded.TorchEngineBase, the code the
gloo tests drive the protocol with). One GPU is enough: nothing here communicates. The NCCL side (csrc/comm.cu) is covered by
tests/test_gpu_sharded.py on boxes with >= 2 GPUs and by bench.py's `sharded_labels_identical` at every N > 1."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


class TorchOnCuda:
    """TorchEngineBase bound to the CUDA device (reference statement of the vectors)."""

    def __new__(cls):
        from radar_point_cloud_tracking_b200 import device as dev
        from radar_point_cloud_tracking_b200.sharded import TorchEngineBase

        class _E(TorchEngineBase):
            device = torch.device("cuda:0")

            def expand_frame_times(self, frame_off, frame_ids, n):
                return dev.expand_frame_times(frame_off, frame_ids, n)
        return _E()


@pytest.fixture(scope="module")
def engines():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(0)
    from radar_point_cloud_tracking_b200.sharded import CudaEngine
    return CudaEngine(0), TorchOnCuda()


def test_pack_stats_and_layout(engines):
    cuda, ref = engines
    rng = np.random.default_rng(3)
    d = torch.device("cuda:0")
    for F, hh in ((40, 2), (7, 5), (5, 5), (64, 0), (1, 1)):
        per = rng.integers(0, 50, F)
        per[rng.random(F) < 0.3] = 0
        off = torch.from_numpy(np.concatenate([[0], np.cumsum(per)]).astype(np.int64)).to(d)
        b4 = torch.tensor([-3.5, 200.25, -17.0, 99.0], dtype=torch.float32, device=d)
        assert torch.equal(cuda.pack_stats(off, b4, 12345), ref.pack_stats(off, b4, 12345))
        ids = np.arange(1000, 1000 + 3 * F, 3)
        assert torch.equal(cuda.pack_layout(off, ids, F, hh), ref.pack_layout(off, ids, F, hh))


def test_local_index_and_pack_keys(engines):
    cuda, ref = engines
    rng = np.random.default_rng(4)
    d = torch.device("cuda:0")
    for nl, n_own, nr in ((300, 5000, 450), (0, 4000, 120), (70, 2500, 0), (0, 900, 0), (0, 0, 0)):
        n_loc = nl + n_own + nr
        frames = [c for c in (3 if nl else 0, 11, 2 if nr else 0)]
        cnts = []
        for total, k in zip((nl, n_own, nr), frames):
            if k:
                cut = np.sort(rng.integers(0, total + 1, k - 1))
                cnts.append(np.diff(np.concatenate([[0], cut, [total]])))
        cnt = np.concatenate(cnts) if cnts else np.zeros(0, np.int64)
        head = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        all_ids = np.arange(50, 50 + len(cnt)).astype(np.float32)
        a = cuda.local_index(head, all_ids, nl, n_own, nr, 10_000, 20_000, 90_000)
        b = ref.local_index(head, all_ids, nl, n_own, nr, 10_000, 20_000, 90_000)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        gidx = a[1]
        # component keys: -1 for non-core points, else the global index of some core point (roots: their own)
        key = torch.full((n_loc,), -1, dtype=torch.int64, device=d)
        if n_loc:
            core = torch.from_numpy(rng.random(n_loc) < 0.6).to(d)
            roots = torch.from_numpy(rng.random(n_loc) < 0.05).to(d) & core
            if not bool(roots.any()) and bool(core.any()):
                roots[torch.nonzero(core)[0]] = True
            root_idx = gidx[roots]
            pick = torch.from_numpy(rng.integers(0, max(int(roots.sum()), 1), n_loc)).to(d)
            key = torch.where(core, root_idx[pick] if len(root_idx) else key, key)
            key[roots] = gidx[roots]
        lo_end, hi_start = n_own // 5, n_own - n_own // 7
        zones = ((0, nl), (nl, nl + lo_end), (nl + hi_start, nl + n_own), (nl + n_own, n_loc))
        for cap_k in (8192, 37, 0):
            got = cuda.pack_keys(key, gidx, zones, n_loc, cap_k)
            want = ref.pack_keys(key, gidx, zones, n_loc, cap_k)
            assert torch.equal(got[:9], want[:9])                           # the five running counts (true sizes), zone lengths
            n_fit = min(int(want[4]), cap_k)
            assert torch.equal(got[9:9 + n_fit], want[9:9 + n_fit])                              # run keys, then root keys
            assert torch.equal(got[9 + cap_k:9 + cap_k + n_fit], want[9 + cap_k:9 + cap_k + n_fit])  # where each run starts
