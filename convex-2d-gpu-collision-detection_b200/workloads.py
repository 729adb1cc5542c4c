"""Synthetic inputs of the BASELINE.json configurations, numpy only.

These build *inputs* (pair descriptors, normal banks); no collision arithmetic happens here.
Distributions follow the reference's dataset front-end:
  poses      w,h ~ U(0.1,5), theta ~ U(0,2pi)             generate_dataset.cu:56-57,319-330
  variances  x,y,theta ~ U(0,0.3), w,h = 0                generate_dataset.cu:54-55,285-300
  positions  ring prior around the obstacle               generate_dataset.cu:211-216
  robot      4.07 x 1.74                                  generate_dataset.cu:60-61
"""
from __future__ import annotations

import numpy as np

from . import PAIR_DTYPE, pairs_from_columns

ROBOT_W, ROBOT_H = 4.07, 1.74


def cfg1_rect_pairs(n: int = 10_000, seed: int = 1):
    """n random rectangle pairs as explicit corner sets ([n,8] each), built in float64 then rounded."""
    rng = np.random.default_rng(seed)

    def rects():
        w, h = rng.uniform(0.1, 5, n), rng.uniform(0.1, 5, n)
        th = rng.uniform(0, 2 * np.pi, n)
        cx, cy = rng.uniform(-6, 6, n), rng.uniform(-6, 6, n)
        bx = np.stack([-w / 2, w / 2, w / 2, -w / 2], 1)
        by = np.stack([-h / 2, -h / 2, h / 2, h / 2], 1)
        c, s = np.cos(th)[:, None], np.sin(th)[:, None]
        x = c * bx - s * by + cx[:, None]
        y = s * bx + c * by + cy[:, None]
        return np.stack([x, y], 2).reshape(n, 8).astype(np.float32)

    return rects(), rects()


def cfg2_pair() -> np.ndarray:
    """One pair: obstacle 2.3 x 1.1, robot at (3.1, 1.9, theta 0.7), variances (0.2, 0.1, 0.15, 0, 0)."""
    return pairs_from_columns(3.1, 1.9, 0.7, 2.3, 1.1, np.sqrt(0.2), np.sqrt(0.1), np.sqrt(0.15))


def normal_bank(n: int, ndof: int = 5, seed: int = 2) -> np.ndarray:
    """[ndof, n] float32 standard normals (SoA planes x, y, theta[, w, h])."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((ndof, n), dtype=np.float32)


def dataset_pairs(n: int, seed: int = 3, shape_variance: bool = False, spread: float = 4.0,
                  max_variance: float = 0.3) -> np.ndarray:
    """n pairs drawn like one generate_dataset batch (cfg 3)."""
    rng = np.random.default_rng(seed)
    ow, oh = rng.uniform(0.1, 5, n), rng.uniform(0.1, 5, n)
    rtheta = rng.uniform(0, 2 * np.pi, n)
    var = rng.uniform(0, max_variance, (n, 5))
    if not shape_variance:
        var[:, 3:] = 0.0
    sd = np.sqrt(var)
    r_offset = (ROBOT_W + ROBOT_H) / 4
    ang = rng.uniform(0, 1, n) * 2 * np.pi
    shift = rng.standard_normal(n) * ((sd[:, 1] + sd[:, 0]) / 2) * spread
    rx = np.cos(ang) * ((ow / 2 + r_offset + 2.35 + sd[:, 0]) + shift)
    ry = np.sin(ang) * ((oh / 2 + r_offset + 2.35 + sd[:, 1]) + shift)
    return pairs_from_columns(rx, ry, rtheta, ow, oh, sd[:, 0], sd[:, 1], sd[:, 2], sd[:, 3], sd[:, 4])


def variance_sweep_pairs(n_pairs: int = 10_000, seed: int = 5) -> np.ndarray:
    """cfg 5: every pair of `dataset_pairs` x a 4x4x4 covariance grid -> n_pairs*64 rows."""
    base = dataset_pairs(n_pairs, seed)
    grid = np.array([0.01, 0.05, 0.15, 0.3])
    vx, vy, vt = np.meshgrid(grid, grid, grid, indexing="ij")
    sdx, sdy, sdt = np.sqrt(vx.ravel()), np.sqrt(vy.ravel()), np.sqrt(vt.ravel())
    out = np.repeat(base, 64)
    out["sd_x"] = np.tile(sdx, n_pairs)
    out["sd_y"] = np.tile(sdy, n_pairs)
    out["sd_theta"] = np.tile(sdt, n_pairs)
    out["sd_w"] = 0.0
    out["sd_h"] = 0.0
    return np.ascontiguousarray(out, dtype=PAIR_DTYPE)


def reference_tables(pairs: np.ndarray):
    """Re-expresses direct pairs in the reference's indirect layout (one pose / std-dev row per pair):
    robot_base[8], poses[n,3], std_devs[n,5], pose_idxs[n], sd_idxs[n], positions[n,2]."""
    n = pairs.size
    rw, rh = float(pairs["rw"][0]), float(pairs["rh"][0])
    robot_base = np.array([-rw / 2, -rh / 2, rw / 2, -rh / 2, rw / 2, rh / 2, -rw / 2, rh / 2], np.float32)
    poses = np.stack([pairs["ow"], pairs["oh"], pairs["rtheta"]], 1).astype(np.float32)
    sds = np.stack([pairs["sd_x"], pairs["sd_y"], pairs["sd_theta"], pairs["sd_w"], pairs["sd_h"]], 1).astype(np.float32)
    idx = np.arange(n, dtype=np.float32)
    pos = np.stack([pairs["rx"], pairs["ry"]], 1).astype(np.float32)
    return robot_base, poses, sds, idx, idx.copy(), pos
