"""radar-b200: B200-native (sm_100a) per-frame detection hot path for radar point-cloud tracking.

Drop-in for the hot functions of ``4_temporal_object_tracker.py`` / ``radar_pipeline``:

* :mod:`.tracker`     ``load_radar_csv``, ``build_frame``, ``build_occupancy_grid``,
  ``identify_land_cells``, ``filter_land_from_frame``, ``st_dbscan`` and :func:`tracker.install`
* :mod:`.clustering`  flat-label ``st_dbscan(coords, times, ...)``
* :mod:`.transforms`  ``polar_to_cartesian``, ``sweep_to_point_cloud``
* :mod:`.fusion`      concat / grid-max gain fusion
* :mod:`.pipeline`    device-resident batch pipeline (what ``bench.py`` times)
* :mod:`.sharded`     time-sharded multi-GPU driver (one process per GPU, NCCL)

All compute is hand-written CUDA behind the C ABI of ``include/radarb200.h``
(``libradarb200.so``); there is no CPU fallback.
"""
from ._lib import RadarB200Error, load as load_library  # noqa: F401

__all__ = ["RadarB200Error", "load_library"]
__version__ = "0.1.0"
