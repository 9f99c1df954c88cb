"""GPU CSV ingest (rb_csv_parse_sweep, SURVEY section 8 f rank 2) against pandas - the reference's own parser
(4_temporal_object_tracker.py:192) - on the same files, including the shapes that must be handed back to pandas."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from radar_point_cloud_tracking_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

SPEC = dict(seed=77, frames=2, spokes=96, bins=160, clutter_p=0.02, land_blobs=1, buoys=2, boats=2)


@pytest.fixture(scope="module")
def trk():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(0)
    from radar_point_cloud_tracking_b200 import tracker
    return tracker


def _same(trk, path, E, expect_device=True):
    want = trk.read_sweep_csv(path, E)
    got = trk.read_sweep_csv_device(path, E)
    if want is None:
        assert got is None
        return None
    a0, s0, e0, g0 = want
    a1, s1, e1, g1 = got
    assert isinstance(e1, torch.Tensor) == expect_device          # device grammar used, or handed back to pandas
    e1 = e1.cpu().numpy().astype(np.float32) if isinstance(e1, torch.Tensor) else e1
    assert g0 == g1 and a0.dtype == a1.dtype and s0.dtype == s1.dtype
    assert np.array_equal(a0, a1, equal_nan=True) and np.array_equal(s0, s1, equal_nan=True)      # (pandas leaves NaN in a broken row)
    assert e0.shape == e1.shape and np.array_equal(e0, e1, equal_nan=True)
    return got


def _lines(path):
    return path.read_text().split("\n")


def test_csv_tree_matches_pandas(trk, tmp_path):
    spec = syn.SweepSpec(**SPEC)
    frames = syn.write_csv_tree(spec, tmp_path)
    for entry in frames:
        for path in entry.values():
            _same(trk, path, spec.bins)


def test_csv_variants(trk, tmp_path):
    spec = syn.SweepSpec(**SPEC)
    path = next(iter(syn.write_csv_tree(spec, tmp_path)[0].values()))
    rows = _lines(path)
    assert rows[-1] == ""                                          # the writer ends the file with a newline
    E = spec.bins

    def variant(name, new_rows, sep="\n"):
        p = tmp_path / name
        p.write_bytes(sep.join(new_rows).encode())
        return p

    # CRLF line ends; no newline at the end of the file
    _same(trk, variant("crlf.csv", rows, "\r\n"), E)
    _same(trk, variant("open_tail.csv", rows[:-1]), E)
    # empty echo fields (NaN -> 0), a decimal Scale, leading zeros
    r = rows[3].split(",")
    r[5], r[5 + E - 1], r[40] = "", "", "007"
    r[1] = "926.5"
    _same(trk, variant("empties.csv", rows[:3] + [",".join(r)] + rows[4:]), E)
    # header only / empty file
    assert trk.read_sweep_csv_device(variant("header_only.csv", [rows[0], ""]), E) is None
    assert _same(trk, variant("empty.csv", [""]), E) is None
    # outside the device grammar: the file goes to pandas whole and behaves as in the reference
    for name, field in (("decimal_echo.csv", "12.5"), ("big_echo.csv", "300"), ("negative_echo.csv", "-3"), ("spaced.csv", " 7")):
        r = rows[5].split(",")
        r[20] = field
        _same(trk, variant(name, rows[:5] + [",".join(r)] + rows[6:]), E, expect_device=False)
    # carriage returns that are not part of "\r\n" end a line for pandas: a CR-only file and a stray CR inside a row
    _same(trk, variant("cr_only.csv", rows, "\r"), E, expect_device=False)
    r = rows[6].split(",")
    r[2] = "3\r"
    _same(trk, variant("stray_cr.csv", rows[:6] + [",".join(r)] + rows[7:]), E, expect_device=False)
    _same(trk, variant("blank_line.csv", rows[:4] + [""] + rows[4:]), E, expect_device=False)
    _same(trk, variant("short_row.csv", rows[:4] + [",".join(rows[4].split(",")[:-3])] + rows[5:]), E, expect_device=False)


def test_load_radar_csv_uses_the_device_parser_and_matches(trk, tmp_path):
    """The reference-facing function end to end: points from the device-parsed file == points from the pandas-parsed one."""
    spec = syn.SweepSpec(**SPEC)
    path = next(iter(syn.write_csv_tree(spec, tmp_path)[0].values()))

    class Cfg:
        NUM_ECHO_COLUMNS, INTENSITY_THRESHOLD, POINT_STRIDE = spec.bins, 10.0, 4
    x, y, z, gain = trk.load_radar_csv(path, _cfg=Cfg)
    angle, scale, echo, g = trk.read_sweep_csv(path, spec.bins)
    (x0, y0, z0), = trk._points_from_sweeps([(angle, scale, echo, g)], 10.0, 4)
    assert gain == g and len(x) > 0
    assert np.array_equal(x, x0) and np.array_equal(y, y0) and np.array_equal(z, z0)
