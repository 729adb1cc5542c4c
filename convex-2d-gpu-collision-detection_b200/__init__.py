"""satmc-b200: B200-native Monte Carlo SAT collision probability (Python host mirror).

Thin ctypes binding over the C ABI in ``include/satmc.h`` (``libsatmc.so``, hand-written CUDA for
sm_100a).  The reference (beautifulv0id/Convex-2D-GPU-Collision-Detection) has no Python API; its
host side is three CUDA C++ programs, so the real host layer lives in ``host/`` (C++).  This module
exists for the tests, ``bench.py`` and multi-GPU plumbing (``torch.distributed``): PyTorch supplies
device memory and streams, nothing else.

There is NO CPU fallback: importing works anywhere (so the C-ABI symbol test can run without a
GPU) but creating a :class:`Context` raises unless a sm_100 device is present, and a missing
``libsatmc.so`` raises at import of the library handle.

The package directory name contains hyphens; import it with
``importlib.import_module("convex-2d-gpu-collision-detection_b200")`` (see ``tests/conftest.py``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
#: SATMC_LIB selects another build of the same sources (the -DSATMC_DEBUG library with device-side bounds asserts)
LIB_PATH = os.environ.get("SATMC_LIB") or os.path.join(_HERE, "libsatmc.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

SATMC_ACCUMULATE = 0x1
SATMC_EXACT_ONLY = 0x2
SHARD_BY_PAIR, SHARD_BY_SAMPLE_RANGE, SHARD_INTERLEAVED = 0, 1, 2
UNIQUE_ID_BYTES = 128

#: numpy dtype of ``satmc_pair`` (include/satmc.h) -- 12 packed float32 = 48 bytes
PAIR_DTYPE = np.dtype([(n, "<f4") for n in
                       ("rx", "ry", "rtheta", "rw", "rh", "ow", "oh", "sd_x", "sd_y", "sd_theta", "sd_w", "sd_h")])
PAIR_FIELDS = PAIR_DTYPE.names

#: numpy dtype of ``satmc_poly_pair`` -- 160 bytes
POLY_MAX = 8
POLY_PAIR_DTYPE = np.dtype([("rx", "<f4"), ("ry", "<f4"), ("rtheta", "<f4"), ("sd_x", "<f4"), ("sd_y", "<f4"), ("sd_theta", "<f4"),
                            ("n_robot", "<u4"), ("n_obstacle", "<u4"), ("robot", "<f4", (2 * POLY_MAX,)),
                            ("obstacle", "<f4", (2 * POLY_MAX,))])

#: every symbol include/satmc.h declares (checked against the built library by tests/test_abi.py)
ABI_SYMBOLS = (
    "satmc_create", "satmc_destroy", "satmc_synchronize", "satmc_last_error", "satmc_version",
    "satmc_launch_count", "satmc_set_profiling", "satmc_last_kernel_ms",
    "satmc_count_fused", "satmc_count_fused_sweep", "satmc_count_streamed", "satmc_decide_streamed", "satmc_fused_normals",
    "satmc_philox_blocks", "satmc_sat_corners", "satmc_exact_evals", "satmc_screen_debug", "satmc_plan_debug",
    "satmc_count_fused_polygons", "satmc_count_streamed_polygons",
    "satmc_mc_step", "satmc_write_collision_probability", "satmc_adaptive_run", "satmc_sample_positions",
    "satmc_device_alloc", "satmc_device_free", "satmc_upload", "satmc_download", "satmc_download_async",
    "satmc_count_fused_host", "satmc_count_streamed_host", "satmc_collision_probability_host",
    "satmc_host_alloc", "satmc_host_free",
    "satmc_shard_range", "satmc_group_create", "satmc_group_unique_id", "satmc_group_create_rank", "satmc_group_destroy",
    "satmc_group_synchronize", "satmc_group_world", "satmc_group_local_count", "satmc_group_rank", "satmc_group_context",
    "satmc_group_last_error", "satmc_group_nccl_version", "satmc_group_hits_capacity", "satmc_group_count_fused",
    "satmc_group_count_fused_host", "satmc_group_set_timing", "satmc_group_last_times", "satmc_group_set_tables",
    "satmc_group_adaptive_run_host", "satmc_group_last_exchange", "satmc_group_set_peer_reduce",
)


class SatmcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"satmc error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> ctypes.CDLL:
    """Loads ``libsatmc.so`` (built in-tree by ``__graft_entry__.build()`` / ``make``). Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, u64, u32, i32, f32p = c.c_void_p, c.c_uint64, c.c_uint32, c.c_int, c.c_void_p
    sig = {
        "satmc_create": (i32, [i32, vp, c.POINTER(vp)]),
        "satmc_destroy": (i32, [vp]),
        "satmc_synchronize": (i32, [vp]),
        "satmc_last_error": (c.c_char_p, [vp]),
        "satmc_version": (c.c_char_p, []),
        "satmc_launch_count": (u64, [vp]),
        "satmc_set_profiling": (i32, [vp, i32]),
        "satmc_last_kernel_ms": (c.c_float, [vp]),
        "satmc_count_fused": (i32, [vp, vp, u64, u64, u64, u64, u32, vp, u32]),
        "satmc_count_fused_sweep": (i32, [vp, vp, u64, f32p, u32, u64, u64, u64, u32, vp, u32]),
        "satmc_count_streamed": (i32, [vp, vp, u64, f32p, u64, u64, i32, u64, vp, u32]),
        "satmc_decide_streamed": (i32, [vp, vp, f32p, u64, i32, u64, vp, u32]),
        "satmc_fused_normals": (i32, [vp, u64, u32, u64, u64, i32, f32p, u64]),
        "satmc_philox_blocks": (i32, [vp, vp, u64, u32, u32, vp]),
        "satmc_sat_corners": (i32, [vp, f32p, f32p, u64, vp]),
        "satmc_exact_evals": (i32, [vp, c.POINTER(u64), i32]),
        "satmc_screen_debug": (i32, [vp, vp, f32p, u64, i32, u64, f32p, f32p]),
        "satmc_plan_debug": (i32, [vp, i32, u64, u64, c.POINTER(u64), c.POINTER(u64)]),
        "satmc_count_fused_polygons": (i32, [vp, vp, u64, u64, u64, u64, u32, vp, u32]),
        "satmc_count_streamed_polygons": (i32, [vp, vp, u64, f32p, u64, u64, u64, vp, u32]),
        "satmc_mc_step": (i32, [vp, f32p, f32p, u32, f32p, u32, f32p, f32p, f32p, f32p, f32p, f32p, i32, vp,
                                i32, i32, i32, i32, u64, u32]),
        "satmc_write_collision_probability": (i32, [vp, f32p, i32, i32]),
        "satmc_adaptive_run": (i32, [vp, f32p, f32p, u32, f32p, u32, f32p, f32p, f32p, i32, f32p, f32p, i32, i32, i32, i32,
                                     i32, u64, u32, f32p, vp, c.POINTER(i32), c.POINTER(c.c_longlong)]),
        "satmc_sample_positions": (i32, [vp, f32p, u32, f32p, u32, i32, c.c_float, c.c_float, u64, u32, f32p, f32p, f32p]),
        "satmc_device_alloc": (i32, [vp, c.POINTER(vp), c.c_size_t]),
        "satmc_device_free": (i32, [vp, vp]),
        "satmc_upload": (i32, [vp, vp, vp, c.c_size_t]),
        "satmc_download": (i32, [vp, vp, vp, c.c_size_t]),
        "satmc_download_async": (i32, [vp, vp, vp, c.c_size_t]),
        "satmc_count_fused_host": (i32, [vp, vp, u64, u64, u64, u64, u32, vp, u32]),
        "satmc_count_streamed_host": (i32, [vp, vp, u64, f32p, u64, u64, i32, u64, vp, u32]),
        "satmc_collision_probability_host": (i32, [vp, vp, u64, u64, u64, f32p]),
        "satmc_host_alloc": (i32, [c.POINTER(vp), c.c_size_t]),
        "satmc_host_free": (i32, [vp]),
        "satmc_shard_range": (i32, [i32, u64, i32, i32, c.POINTER(u64), c.POINTER(u64)]),
        "satmc_group_create": (i32, [c.POINTER(i32), i32, c.POINTER(vp)]),
        "satmc_group_unique_id": (i32, [vp]),
        "satmc_group_create_rank": (i32, [vp, i32, i32, i32, vp, c.POINTER(vp)]),
        "satmc_group_destroy": (i32, [vp]),
        "satmc_group_synchronize": (i32, [vp]),
        "satmc_group_world": (i32, [vp]),
        "satmc_group_local_count": (i32, [vp]),
        "satmc_group_rank": (i32, [vp, i32]),
        "satmc_group_context": (vp, [vp, i32]),
        "satmc_group_last_error": (c.c_char_p, [vp]),
        "satmc_group_nccl_version": (i32, []),
        "satmc_group_hits_capacity": (u64, [vp, u64]),
        "satmc_group_count_fused": (i32, [vp, c.POINTER(vp), u64, u64, u64, u64, u32, i32, c.POINTER(vp), u32]),
        "satmc_group_count_fused_host": (i32, [vp, vp, u64, u64, u64, u64, u32, i32, vp, u32]),
        "satmc_group_set_timing": (i32, [vp, i32]),
        "satmc_group_last_exchange": (i32, [vp]),
        "satmc_group_set_peer_reduce": (i32, [vp, i32]),
        "satmc_group_last_times": (i32, [vp, c.POINTER(c.c_float), c.POINTER(c.c_float)]),
        "satmc_group_set_tables": (i32, [vp, f32p, f32p, u32, f32p, u32, f32p, f32p, i32]),
        "satmc_group_adaptive_run_host": (i32, [vp, f32p, f32p, f32p, i32, i32, i32, i32, i32, u64, u32, f32p, c.POINTER(i32),
                                                c.POINTER(c.c_longlong)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def make_pairs(n: int) -> np.ndarray:
    """Zero-initialised structured array of ``n`` pairs."""
    return np.zeros(n, dtype=PAIR_DTYPE)


def pairs_from_columns(rx, ry, rtheta, ow, oh, sd_x, sd_y, sd_theta, sd_w=0.0, sd_h=0.0,
                       rw=4.07, rh=1.74) -> np.ndarray:
    """Builds the pair array from broadcastable columns (robot defaults: generate_dataset.cu:60-61)."""
    cols = np.broadcast_arrays(*[np.asarray(a, dtype=np.float32) for a in
                                 (rx, ry, rtheta, rw, rh, ow, oh, sd_x, sd_y, sd_theta, sd_w, sd_h)])
    out = make_pairs(cols[0].size)
    for name, col in zip(PAIR_FIELDS, cols):
        out[name] = col.ravel()
    return out


def make_poly_pairs(robots, obstacles, rx, ry, rtheta, sd_x, sd_y, sd_theta) -> np.ndarray:
    """Polygon pair array: `robots` / `obstacles` are sequences of [k,2] CCW vertex arrays (k <= 8), the rest broadcast."""
    n = len(robots)
    out = np.zeros(n, dtype=POLY_PAIR_DTYPE)
    cols = np.broadcast_arrays(*[np.asarray(a, dtype=np.float32) for a in (rx, ry, rtheta, sd_x, sd_y, sd_theta)], np.zeros(n, np.float32))
    for name, col in zip(("rx", "ry", "rtheta", "sd_x", "sd_y", "sd_theta"), cols):
        out[name] = col
    for i, (r, o) in enumerate(zip(robots, obstacles)):
        r = np.asarray(r, np.float32).reshape(-1, 2); o = np.asarray(o, np.float32).reshape(-1, 2)
        if not (1 <= len(r) <= POLY_MAX and 1 <= len(o) <= POLY_MAX):
            raise ValueError("polygons need 1..8 vertices")
        out["n_robot"][i] = len(r); out["n_obstacle"][i] = len(o)
        out["robot"][i, :2 * len(r)] = r.ravel(); out["obstacle"][i, :2 * len(o)] = o.ravel()
    return out


def _ptr(x) -> int:
    """Raw address of a torch CUDA tensor, numpy array, or int."""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()          # torch.Tensor


class Context:
    """One satmc context = one GPU + one stream (``satmc_create``)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._lib = load_library()
        h = ctypes.c_void_p()
        rc = self._lib.satmc_create(int(device), ctypes.c_void_p(stream or 0), ctypes.byref(h))
        if rc != 0:
            raise SatmcError(rc, self._lib.satmc_last_error(None).decode())
        self._h = h
        self.device = int(device)

    # -- plumbing -----------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise SatmcError(rc, self._lib.satmc_last_error(self._h).decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.satmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def synchronize(self) -> None:
        self._check(self._lib.satmc_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.satmc_launch_count(self._h))

    def set_profiling(self, on: bool) -> None:
        self._check(self._lib.satmc_set_profiling(self._h, int(bool(on))))

    def last_kernel_ms(self) -> float:
        return float(self._lib.satmc_last_kernel_ms(self._h))

    def exact_evals(self, reset: bool = False) -> int:
        v = ctypes.c_uint64()
        self._check(self._lib.satmc_exact_evals(self._h, ctypes.byref(v), int(reset)))
        return int(v.value)

    def plan_debug(self, kind: int, n_pairs: int, n_samples: int):
        """(samples per work item, items per pair) the planner picks; kind 0 fused, 1 streamed, 2 polygons, 3 sweep."""
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self._lib.satmc_plan_debug(self._h, kind, n_pairs, n_samples, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    # -- device-pointer entry points (torch CUDA tensors or raw addresses) -----------------------
    def count_fused(self, d_pairs, n_pairs, n_samples, seed, d_hits, sample_offset=0, pair_id_offset=0, flags=0):
        self._check(self._lib.satmc_count_fused(self._h, _ptr(d_pairs), n_pairs, n_samples, seed, sample_offset,
                                                pair_id_offset, _ptr(d_hits), flags))

    def count_fused_sweep(self, d_pairs, n_pairs, sigmas, n_cov, n_samples, seed, d_hits, sample_offset=0, pair_id_offset=0, flags=0):
        """`sigmas`: HOST array of n_cov x (sd_x, sd_y, sd_theta)."""
        sig = np.ascontiguousarray(sigmas, dtype=np.float32).ravel()
        if sig.size < 3 * n_cov:
            raise ValueError("sigmas needs 3 * n_cov values")
        self._check(self._lib.satmc_count_fused_sweep(self._h, _ptr(d_pairs), n_pairs, sig.ctypes.data, n_cov, n_samples, seed,
                                                      sample_offset, pair_id_offset, _ptr(d_hits), flags))

    def count_streamed(self, d_pairs, n_pairs, d_z, ldz, ndof, n_samples, d_hits, z_pair_stride=0, flags=0):
        self._check(self._lib.satmc_count_streamed(self._h, _ptr(d_pairs), n_pairs, _ptr(d_z), ldz, z_pair_stride,
                                                   ndof, n_samples, _ptr(d_hits), flags))

    def count_fused_polygons(self, d_pairs, n_pairs, n_samples, seed, d_hits, sample_offset=0, pair_id_offset=0, flags=0):
        self._check(self._lib.satmc_count_fused_polygons(self._h, _ptr(d_pairs), n_pairs, n_samples, seed, sample_offset,
                                                         pair_id_offset, _ptr(d_hits), flags))

    def count_streamed_polygons(self, d_pairs, n_pairs, d_z, ldz, n_samples, d_hits, z_pair_stride=0, flags=0):
        self._check(self._lib.satmc_count_streamed_polygons(self._h, _ptr(d_pairs), n_pairs, _ptr(d_z), ldz, z_pair_stride,
                                                            n_samples, _ptr(d_hits), flags))

    def decide_streamed(self, d_pair, d_z, ldz, ndof, n_samples, d_out, flags=0):
        self._check(self._lib.satmc_decide_streamed(self._h, _ptr(d_pair), _ptr(d_z), ldz, ndof, n_samples,
                                                    _ptr(d_out), flags))

    def screen_debug(self, d_pair, d_z, ldz, ndof, n_samples, d_m_out, d_eps_out):
        self._check(self._lib.satmc_screen_debug(self._h, _ptr(d_pair), _ptr(d_z), ldz, ndof, n_samples, _ptr(d_m_out),
                                                 _ptr(d_eps_out)))

    def fused_normals(self, seed, pair_id, sample_offset, n, ndof, d_z, ldz):
        self._check(self._lib.satmc_fused_normals(self._h, seed, pair_id, sample_offset, n, ndof, _ptr(d_z), ldz))

    def philox_blocks(self, d_ctr, n, key0, key1, d_out):
        self._check(self._lib.satmc_philox_blocks(self._h, _ptr(d_ctr), n, key0, key1, _ptr(d_out)))

    def sat_corners(self, d_r1, d_r2, n, d_out):
        self._check(self._lib.satmc_sat_corners(self._h, _ptr(d_r1), _ptr(d_r2), n, _ptr(d_out)))

    def mc_step(self, d_robot_base, d_poses, n_poses, d_std_devs, n_std, d_pose_idxs, d_std_dev_idxs, d_positions,
                d_cps, d_accuracy_bins, d_bin_accuracy, n_accuracy_bins, d_done, iteration, n_samples, n_batch,
                num_left, seed, stream_id_offset=0):
        self._check(self._lib.satmc_mc_step(self._h, _ptr(d_robot_base), _ptr(d_poses), n_poses, _ptr(d_std_devs),
                                            n_std, _ptr(d_pose_idxs), _ptr(d_std_dev_idxs), _ptr(d_positions),
                                            _ptr(d_cps), _ptr(d_accuracy_bins), _ptr(d_bin_accuracy),
                                            n_accuracy_bins, _ptr(d_done), iteration, n_samples, n_batch, num_left,
                                            seed, stream_id_offset))

    def adaptive_run(self, d_robot_base, d_poses, n_poses, d_std_devs, n_std, d_pose_idxs, d_std_dev_idxs, d_positions,
                     n_pairs, d_accuracy_bins, d_bin_accuracy, n_accuracy_bins, max_samples, n_batch_small, switch_at,
                     n_batch_large, seed, d_cp_out, d_n_samples_out=None, stream_id_offset=0):
        """Returns (iterations, samples_drawn)."""
        it, drawn = ctypes.c_int(0), ctypes.c_longlong(0)
        self._check(self._lib.satmc_adaptive_run(self._h, _ptr(d_robot_base), _ptr(d_poses), n_poses, _ptr(d_std_devs), n_std,
                                                 _ptr(d_pose_idxs), _ptr(d_std_dev_idxs), _ptr(d_positions), n_pairs,
                                                 _ptr(d_accuracy_bins), _ptr(d_bin_accuracy), n_accuracy_bins, max_samples,
                                                 n_batch_small, switch_at, n_batch_large, seed, stream_id_offset,
                                                 _ptr(d_cp_out), _ptr(d_n_samples_out), ctypes.byref(it), ctypes.byref(drawn)))
        return int(it.value), int(drawn.value)

    def sample_positions(self, d_poses, n_poses, d_std_devs, n_std, n, r_offset, spread, seed, d_positions, d_pose_idxs,
                         d_std_dev_idxs, stream_id_offset=0):
        self._check(self._lib.satmc_sample_positions(self._h, _ptr(d_poses), n_poses, _ptr(d_std_devs), n_std, n, r_offset,
                                                     spread, seed, stream_id_offset, _ptr(d_positions), _ptr(d_pose_idxs),
                                                     _ptr(d_std_dev_idxs)))

    def write_collision_probability(self, d_counts, n_done, n_samples):
        self._check(self._lib.satmc_write_collision_probability(self._h, _ptr(d_counts), n_done, n_samples))

    # -- host-buffer entry points (numpy) -----------------------------------------------------------
    def count_fused_host(self, pairs: np.ndarray, n_samples: int, seed: int, sample_offset: int = 0,
                         pair_id_offset: int = 0, flags: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        hits = out if out is not None else np.zeros(pairs.size, dtype=np.uint64)
        self._check(self._lib.satmc_count_fused_host(self._h, pairs.ctypes.data, pairs.size, n_samples, seed,
                                                     sample_offset, pair_id_offset, hits.ctypes.data, flags))
        return hits

    def count_streamed_host(self, pairs: np.ndarray, z: np.ndarray, n_samples: Optional[int] = None,
                            z_pair_stride: int = 0, flags: int = 0) -> np.ndarray:
        """``z`` is ``[ndof, ldz]`` float32 (SoA planes)."""
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        z = np.ascontiguousarray(z, dtype=np.float32)
        ndof, ldz = z.shape
        if n_samples is None:
            n_samples = ldz if z_pair_stride == 0 else z_pair_stride
        hits = np.zeros(pairs.size, dtype=np.uint64)
        self._check(self._lib.satmc_count_streamed_host(self._h, pairs.ctypes.data, pairs.size, z.ctypes.data, ldz,
                                                        z_pair_stride, ndof, n_samples, hits.ctypes.data, flags))
        return hits

    def collision_probability_host(self, pairs: np.ndarray, n_samples: int, seed: int) -> np.ndarray:
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        cp = np.zeros(pairs.size, dtype=np.float32)
        self._check(self._lib.satmc_collision_probability_host(self._h, pairs.ctypes.data, pairs.size, n_samples, seed,
                                                               cp.ctypes.data))
        return cp


def shard_range(mode: int, n_units: int, world: int, rank: int):
    """``satmc_shard_range``: the slice of `rank` (host arithmetic, no GPU needed)."""
    lib = load_library()
    lo, hi = ctypes.c_uint64(), ctypes.c_uint64()
    rc = lib.satmc_shard_range(mode, n_units, world, rank, ctypes.byref(lo), ctypes.byref(hi))
    if rc != 0:
        raise SatmcError(rc, lib.satmc_last_error(None).decode())
    return int(lo.value), int(hi.value)


class Group:
    """``satmc_group``: the path sharded over several GPUs, the collective (NCCL) inside the C library.

    ``Group(devices=[0, 1, ...])``: this process drives all the devices (``satmc_group_create``).
    ``Group.from_rank(unique_id, world, rank, device, stream)``: one process per GPU (``satmc_group_create_rank``);
    ``Group.unique_id()`` on rank 0 yields the 128 bytes the launcher has to distribute."""

    def __init__(self, devices=None, _handle=None):
        self._lib = load_library()
        if _handle is not None:
            self._h = _handle
            return
        h = ctypes.c_void_p()
        if devices is None:
            raise ValueError("devices required")
        arr = (ctypes.c_int * len(devices))(*devices)
        rc = self._lib.satmc_group_create(arr, len(devices), ctypes.byref(h))
        if rc != 0:
            raise SatmcError(rc, self._lib.satmc_group_last_error(None).decode())
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        lib = load_library()
        buf = ctypes.create_string_buffer(UNIQUE_ID_BYTES)
        rc = lib.satmc_group_unique_id(buf)
        if rc != 0:
            raise SatmcError(rc, lib.satmc_group_last_error(None).decode())
        return buf.raw

    @classmethod
    def from_rank(cls, unique_id: Optional[bytes], world: int, rank: int, device: int, stream: Optional[int] = None):
        lib = load_library()
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(unique_id, UNIQUE_ID_BYTES) if unique_id is not None else None
        rc = lib.satmc_group_create_rank(buf, world, rank, device, ctypes.c_void_p(stream or 0), ctypes.byref(h))
        if rc != 0:
            raise SatmcError(rc, lib.satmc_group_last_error(None).decode())
        return cls(_handle=h)

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise SatmcError(rc, self._lib.satmc_group_last_error(self._h).decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.satmc_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def world(self) -> int:
        return int(self._lib.satmc_group_world(self._h))

    @property
    def local_count(self) -> int:
        return int(self._lib.satmc_group_local_count(self._h))

    def rank(self, local: int = 0) -> int:
        return int(self._lib.satmc_group_rank(self._h, local))

    def launch_count(self) -> int:
        return sum(int(self._lib.satmc_launch_count(self._lib.satmc_group_context(self._h, l))) for l in range(self.local_count))

    def synchronize(self) -> None:
        self._check(self._lib.satmc_group_synchronize(self._h))

    def hits_capacity(self, n_pairs: int) -> int:
        return int(self._lib.satmc_group_hits_capacity(self._h, n_pairs))

    def set_peer_reduce(self, on) -> None:
        """Force on (True) / off (False) / automatic (None) the collective-free sample-range reduction over peer memory
        (count_fused_host, one process)."""
        self._check(self._lib.satmc_group_set_peer_reduce(self._h, -1 if on is None else int(bool(on))))

    def last_exchange(self) -> str:
        return {0: "none", 1: "nccl", 2: "peer_atomics"}[int(self._lib.satmc_group_last_exchange(self._h))]

    def set_timing(self, on: bool) -> None:
        self._check(self._lib.satmc_group_set_timing(self._h, int(bool(on))))

    def last_times(self):
        """(kernel_ms, collective_ms) of local device 0 for the last timed count_fused."""
        k, c = ctypes.c_float(), ctypes.c_float()
        self._check(self._lib.satmc_group_last_times(self._h, ctypes.byref(k), ctypes.byref(c)))
        return float(k.value), float(c.value)

    def count_fused(self, d_pairs, n_pairs, n_samples, seed, shard_mode, d_hits, sample_offset=0, pair_id_offset=0, flags=0):
        """d_pairs / d_hits: one tensor (or address) per local device."""
        if not isinstance(d_pairs, (list, tuple)):
            d_pairs, d_hits = [d_pairs], [d_hits]
        n = len(d_pairs)
        pp = (ctypes.c_void_p * n)(*[_ptr(x) for x in d_pairs])
        hh = (ctypes.c_void_p * n)(*[_ptr(x) for x in d_hits])
        self._check(self._lib.satmc_group_count_fused(self._h, pp, n_pairs, n_samples, seed, sample_offset, pair_id_offset,
                                                      shard_mode, hh, flags))

    def count_fused_host(self, pairs: np.ndarray, n_samples: int, seed: int, shard_mode: int, sample_offset: int = 0,
                         pair_id_offset: int = 0, flags: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        hits = out if out is not None else np.zeros(pairs.size, dtype=np.uint64)
        self._check(self._lib.satmc_group_count_fused_host(self._h, pairs.ctypes.data, pairs.size, n_samples, seed, sample_offset,
                                                           pair_id_offset, shard_mode, hits.ctypes.data, flags))
        return hits

    def set_tables(self, robot_base, poses, std_devs, accuracy_bins, bin_accuracy) -> None:
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        rb, po, sd, bi, ac = f(robot_base), f(poses), f(std_devs), f(accuracy_bins), f(bin_accuracy)
        self._check(self._lib.satmc_group_set_tables(self._h, rb.ctypes.data, po.ctypes.data, po.size // 3, sd.ctypes.data, sd.size // 5,
                                                     bi.ctypes.data, ac.ctypes.data, bi.size))

    def adaptive_run_host(self, pose_idxs, std_dev_idxs, positions, max_samples, n_batch_small, switch_at, n_batch_large, seed,
                          stream_id_offset=0):
        """Returns (cp, iterations, samples_drawn_by_local_devices)."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        pi, si, pos = f(pose_idxs), f(std_dev_idxs), f(positions)
        cp = np.zeros(pi.size, dtype=np.float32)
        it, drawn = ctypes.c_int(0), ctypes.c_longlong(0)
        self._check(self._lib.satmc_group_adaptive_run_host(self._h, pi.ctypes.data, si.ctypes.data, pos.ctypes.data, pi.size,
                                                            max_samples, n_batch_small, switch_at, n_batch_large, seed, stream_id_offset,
                                                            cp.ctypes.data, ctypes.byref(it), ctypes.byref(drawn)))
        return cp, int(it.value), int(drawn.value)


def compute_collision_probability(pairs: np.ndarray, n_samples: int, seed: int = 0, device: int = 0) -> np.ndarray:
    """One-shot convenience: collision probability of every pair from ``n_samples`` fused samples."""
    with Context(device) as ctx:
        return ctx.collision_probability_host(pairs, n_samples, seed)
