"""Oracle: the input contract of the path — multi-sweep point loading.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED by outputs of the reference's own
mmdet3d/datasets/pipelines/loading.py (LoadPointsFromFile + LoadPointsFromMultiSweeps with core/points/*.py) run in the
build container through tests/golden/ref_shim.py on synthetic sweep files (tests/golden/ref_sweeps.npz,
tests/test_oracle_vs_reference.py).  numpy restatement, same dtypes as the reference at every step (float32 sweeps,
float64 calibration, results cast back to float32).
"""
import numpy as np


def remove_close(points, radius=1.0):
    """ref: loading.py:824-844: drop points with |x| < radius AND |y| < radius."""
    close = np.logical_and(np.abs(points[:, 0]) < radius, np.abs(points[:, 1]) < radius)
    return points[np.logical_not(close)]


def load_points_from_multi_sweeps(key_points, ts, sweeps, sweeps_num=10, load_dim=5, use_dim=(0, 1, 2, 3, 4),
                                  pad_empty_sweeps=False, remove_close_pts=False, test_mode=False, rng=None):
    """ref: LoadPointsFromMultiSweeps.__call__ (loading.py:846-906).

    key_points [N0, load_dim] float32 (the key frame as LoadPointsFromFile delivers it); ts = results['timestamp'] (s);
    sweeps = list of dict(points [Ni, load_dim] float32 raw, timestamp (us), sensor2lidar_rotation [3,3] float64,
    sensor2lidar_translation [3] float64).  Returns [N, len(use_dim)] float32: key frame (time lag 0, :862) followed by the
    chosen sweeps in order, each rotated + translated into the key frame's LiDAR frame (:886-888; float64 arithmetic,
    stored as float32 after the rotation and again after the translation) with column 4 = ts - sweep_ts/1e6 (:889)."""
    points = np.array(key_points, dtype=np.float32, copy=True)
    points[:, 4] = 0
    out = [points]
    if pad_empty_sweeps and len(sweeps) == 0:                          # :865-870
        for _ in range(sweeps_num):
            out.append(remove_close(points) if remove_close_pts else points)
    else:
        if len(sweeps) <= sweeps_num:                                  # :872-878
            choices = np.arange(len(sweeps))
        elif test_mode:
            choices = np.arange(sweeps_num)
        else:
            choices = (rng or np.random).choice(len(sweeps), sweeps_num, replace=False)
        for idx in choices:
            sweep = sweeps[idx]
            ps = np.copy(sweep['points']).reshape(-1, load_dim)
            if remove_close_pts:
                ps = remove_close(ps)
            sweep_ts = sweep['timestamp'] / 1e6
            ps[:, :3] = ps[:, :3] @ sweep['sensor2lidar_rotation'].T
            ps[:, :3] += sweep['sensor2lidar_translation']
            ps[:, 4] = ts - sweep_ts
            out.append(ps)
    return np.concatenate(out, 0)[:, list(use_dim)]
