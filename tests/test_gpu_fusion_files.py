"""File-level fusion surface of ``5_gain_fusion_ply_builder.py`` (``fusion.install``): a stub module with T5's
configuration globals gets the GPU functions, results against the oracle on a synthetic CSV tree."""
from types import SimpleNamespace

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import numpy_oracle as O
from radar_point_cloud_tracking_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

SPEC = dict(seed=91, frames=1, spokes=64, bins=128, clutter_p=0.03, land_blobs=1, buoys=2, boats=2)


def test_installed_fusion_functions_match_the_oracle(tmp_path):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")
    from radar_point_cloud_tracking_b200 import fusion, plyio

    spec = syn.SweepSpec(**SPEC)
    echo = syn.synth_echo(spec)
    frame_files = syn.write_csv_tree(spec, tmp_path, echo)[0]
    t5 = SimpleNamespace(NUM_ECHO_COLUMNS=spec.bins, INTENSITY_THRESHOLD=5.0, POINT_STRIDE=8)      # T5:52-58 with this tree's width
    fusion.install(t5)
    per_gain = {g: O.sweep_to_points(echo[0, gi], spec.angle_units(), spec.scale(), 5.0, 8) for gi, g in enumerate(spec.gains)}
    pts, gains = O.fuse_concat(per_gain)
    x, y, z, lab = t5.fuse_gains_absolute(frame_files)
    assert np.array_equal(np.column_stack([x, y, z]), pts) and np.array_equal(lab, gains) and lab.dtype == np.int32
    for res in (1.0, 2.5):
        want = O.fuse_max(per_gain, res)
        got = t5.fuse_gains_max(frame_files, res)
        assert all(a.dtype == b.dtype and np.array_equal(a, b) for a, b in zip(got, want))
    t5.POINT_STRIDE = 3                                            # the module's globals are read at call time
    x3, _, _, g3 = t5.load_radar_csv(frame_files[spec.gains[0]])
    assert np.array_equal(x3, O.sweep_to_points(echo[0, 0], spec.angle_units(), spec.scale(), 5.0, 3)[0]) and g3 == spec.gains[0]
    assert t5.write_ply_fast is plyio.write_ply_fast and t5.intensity_to_rgb is plyio.intensity_to_rgb
