"""Golden CSVs of a REAL ``run_pipeline`` run of the unmodified reference (T4:893-1038 -> save_tracking_results
T4:832-886) on the PIPE_SPEC CSV tree: the texts of ``clusters.csv``, ``trajectories.csv`` and
``tracked_objects.csv``. Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden_run_pipeline.py

``tests/test_gpu_reference_install.py`` patches the real T4 module (a copy under ``baseline/_ref/``) with
``tracker.install`` and must reproduce these files byte for byte."""
from __future__ import annotations

import contextlib
import importlib.util
import io
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(HERE))

from radar_point_cloud_tracking_b200 import synthetic as syn  # noqa: E402
from specs import PIPE_SPEC  # noqa: E402

NAMES = ("clusters.csv", "trajectories.csv", "tracked_objects.csv")


def main():
    spec_ = importlib.util.spec_from_file_location("ref_t4", REF / "PointCloudWork" / "4_temporal_object_tracker.py")
    T4 = importlib.util.module_from_spec(spec_)
    sys.modules["ref_t4"] = T4
    spec_.loader.exec_module(T4)
    spec = syn.SweepSpec(**PIPE_SPEC)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        root = Path(tmp) / "data"
        syn.write_csv_tree(spec, root)
        for tag, kw in (("default", {}), ("nofilter_eps6", dict(skip_land_filter=True, eps_space=6.0, eps_time=1.0, min_samples=8))):
            outdir = Path(tmp) / f"out_{tag}"
            with contextlib.redirect_stdout(io.StringIO()):
                T4.run_pipeline(root, outdir, visualize=False, **kw)
            for name in NAMES:
                out[f"{tag}/{name}"] = np.array((outdir / name).read_text())
    np.savez_compressed(HERE / "run_pipeline_csv.npz", **out)
    print({k: len(str(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
