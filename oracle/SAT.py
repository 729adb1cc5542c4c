"""SAT.py -- single rectangle-pair overlap check (BASELINE config 1), CPU only.

TEST INFRASTRUCTURE (oracle).  The reference's README.md:3 describes a SAT.py that is absent from the
repository snapshot; this file stands in for it as the numpy front-end of the oracle:

  collide(r1, r2)        one pair, through the C restatement (bit-exact reference arithmetic)
  collide_many(r1, r2)   [n,8] x [n,8] through the C restatement
  collide_numpy(r1, r2)  the same algorithm vectorised in numpy; float32 FMA is emulated in float64
                         (exact product, one float64 add, round to float32: differs from a true fmaf
                         only by double rounding, which the test suite checks against the C path)

Rectangles are 8 floats x0,y0..x3,y3, corners in order (utils.cu:119-130).
Usage: python oracle/SAT.py x0 y0 ... x3 y3  x0 y0 ... x3 y3
"""
from __future__ import annotations

import sys

import numpy as np


def _oracle():
    from oracle.binding import Oracle
    return Oracle()


def collide(r1, r2) -> int:
    return _oracle().convex_collide(r1, r2)


def collide_many(r1, r2) -> np.ndarray:
    return _oracle().sat_batch(r1, r2)


def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def collide_numpy(r1, r2) -> np.ndarray:
    """utils.cu:159-184 vectorised over n pairs; returns uint8[n]."""
    r1 = np.asarray(r1, np.float32).reshape(-1, 4, 2)
    r2 = np.asarray(r2, np.float32).reshape(-1, 4, 2)
    collide_ = np.ones(r1.shape[0], bool)
    for r in (r1, r2):
        for i in range(4):
            n = r[:, (i + 1) % 4, :] - r[:, i, :]                    # float32 subtract
            n0, n1 = n[:, 0:1], n[:, 1:2]
            p1 = _fma32(n0, r1[:, :, 0], (n1 * r1[:, :, 1]).astype(np.float32))
            p2 = _fma32(n0, r2[:, :, 0], (n1 * r2[:, :, 1]).astype(np.float32))
            sep = (p1.max(1) < p2.min(1)) | (p2.max(1) < p1.min(1))
            collide_ &= ~sep
    return collide_.astype(np.uint8)


if __name__ == "__main__":
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    v = [float(x) for x in sys.argv[1:]]
    if len(v) != 16:
        sys.exit(__doc__)
    print(collide(v[:8], v[8:]))
