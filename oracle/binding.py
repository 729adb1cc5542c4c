"""ctypes bindings for the CPU oracle (libsat_oracle.so) and the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libsat_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_gpu.so")

PAIR_DTYPE = np.dtype([(n, "<f4") for n in
                       ("rx", "ry", "rtheta", "rw", "rh", "ow", "oh", "sd_x", "sd_y", "sd_theta", "sd_w", "sd_h")])


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "sat_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "libsat_oracle.so"])
    return ORACLE_SO


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Oracle:
    """The C restatement (sat_oracle.h)."""

    def __init__(self):
        self.lib = L = C.CDLL(build_oracle())
        fp, u8p, u64p = C.c_void_p, C.c_void_p, C.c_void_p
        L.orc_cuda_sinf.restype = C.c_float; L.orc_cuda_sinf.argtypes = [C.c_float]
        L.orc_cuda_cosf.restype = C.c_float; L.orc_cuda_cosf.argtypes = [C.c_float]
        L.orc_create_rect.argtypes = [fp, C.c_float, C.c_float]
        L.orc_rot_trans_rectangle.argtypes = [fp, C.c_float, C.c_float, C.c_float]
        L.orc_sample_rectangle.argtypes = [fp, fp, fp, fp]
        L.orc_convex_collide.restype = C.c_int; L.orc_convex_collide.argtypes = [fp, fp]
        L.orc_calc_slack.restype = C.c_float; L.orc_calc_slack.argtypes = [C.c_int, C.c_int]
        L.orc_get_bin.restype = C.c_int; L.orc_get_bin.argtypes = [C.c_float, fp, C.c_int]
        L.orc_mc_thread.restype = C.c_int
        L.orc_mc_thread.argtypes = [fp, C.c_float, C.c_float, C.c_float, fp, C.c_float, C.c_float, C.c_int, fp,
                                    C.c_size_t, C.c_int, C.c_int, fp, fp, C.c_int, C.c_void_p]
        L.orc_robot_corners.argtypes = [fp, fp]
        L.orc_sat_batch.argtypes = [fp, fp, C.c_size_t, u8p]
        L.orc_count_streamed.restype = C.c_uint64
        L.orc_count_streamed.argtypes = [fp, fp, C.c_size_t, C.c_int, C.c_size_t, u8p]
        L.orc_count_streamed_batch.argtypes = [fp, C.c_size_t, fp, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, u64p, C.c_int]
        L.orc_poly_count_streamed.restype = C.c_uint64
        L.orc_poly_count_streamed.argtypes = [fp, fp, C.c_size_t, C.c_size_t, u8p]
        L.orc_philox4x32_10.argtypes = [fp, fp, fp]
        L.orc_fused_normals.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, fp]
        L.orc_count_fused_batch.argtypes = [fp, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, u64p, C.c_int]
        L.orc_hardware_threads.restype = C.c_int
        L.orc_affinity_count.restype = C.c_int
        L.orc_sat_batch_mt.argtypes = [fp, fp, C.c_size_t, u8p, C.c_int, C.c_int]
        L.orc_count_streamed_mt.restype = C.c_uint64
        L.orc_count_streamed_mt.argtypes = [fp, fp, C.c_size_t, C.c_int, C.c_size_t, C.c_int]

    # scalar helpers ---------------------------------------------------------------------------
    def cuda_sinf(self, x): return float(self.lib.orc_cuda_sinf(float(x)))
    def cuda_cosf(self, x): return float(self.lib.orc_cuda_cosf(float(x)))
    def calc_slack(self, n, k): return float(self.lib.orc_calc_slack(int(n), int(k)))
    def hardware_threads(self): return int(self.lib.orc_hardware_threads())
    def affinity_count(self): return int(self.lib.orc_affinity_count())

    def sat_batch_mt(self, r1, r2, reps=1, threads=1):
        """convex_collide over all pairs, `reps` times, on `threads` threads (BASELINE.md 4a run C1)."""
        r1, r2 = _f32(r1).reshape(-1, 8), _f32(r2).reshape(-1, 8)
        out = np.zeros(r1.shape[0], np.uint8)
        self.lib.orc_sat_batch_mt(r1.ctypes.data, r2.ctypes.data, r1.shape[0], out.ctypes.data, int(reps), int(threads))
        return out

    def count_streamed_mt(self, pair, z, threads=1):
        """one pair on shared normals, the sample range split over `threads` threads (run C2)."""
        p = np.ascontiguousarray(pair, dtype=PAIR_DTYPE).reshape(1)
        z = _f32(z)
        ndof, ldz = z.shape
        return int(self.lib.orc_count_streamed_mt(p.ctypes.data, z.ctypes.data, ldz, ndof, ldz, int(threads)))

    def get_bin(self, p, bins):
        b = _f32(list(bins) + [0.0])                 # one readable entry past the end (utils.cu:202)
        return int(self.lib.orc_get_bin(float(p), b.ctypes.data, len(bins)))

    def create_rect(self, w, h):
        r = np.zeros(8, np.float32)
        self.lib.orc_create_rect(r.ctypes.data, float(w), float(h))
        return r

    def rot_trans(self, r, dx, dy, dt):
        r = _f32(r).copy()
        self.lib.orc_rot_trans_rectangle(r.ctypes.data, float(dx), float(dy), float(dt))
        return r

    def sample_rectangle(self, r_in, sd, z):
        r_in, sd, z = _f32(r_in), _f32(sd), _f32(z)
        out = np.zeros(8, np.float32)
        self.lib.orc_sample_rectangle(r_in.ctypes.data, out.ctypes.data, sd.ctypes.data, z.ctypes.data)
        return out

    def convex_collide(self, r1, r2):
        r1, r2 = _f32(r1), _f32(r2)
        return int(self.lib.orc_convex_collide(r1.ctypes.data, r2.ctypes.data))

    def robot_corners(self, pair):
        p = np.ascontiguousarray(pair, dtype=PAIR_DTYPE).reshape(1)
        r = np.zeros(8, np.float32)
        self.lib.orc_robot_corners(p.ctypes.data, r.ctypes.data)
        return r

    # batched -----------------------------------------------------------------------------------
    def sat_batch(self, r1, r2):
        r1, r2 = _f32(r1).reshape(-1, 8), _f32(r2).reshape(-1, 8)
        out = np.zeros(r1.shape[0], np.uint8)
        self.lib.orc_sat_batch(r1.ctypes.data, r2.ctypes.data, r1.shape[0], out.ctypes.data)
        return out

    def count_streamed(self, pair, z, n=None, want_decisions=False):
        """z: [ndof, ldz] float32. Returns hits (and decisions)."""
        p = np.ascontiguousarray(pair, dtype=PAIR_DTYPE).reshape(1)
        z = _f32(z)
        ndof, ldz = z.shape
        n = ldz if n is None else n
        dec = np.zeros(n, np.uint8) if want_decisions else None
        h = self.lib.orc_count_streamed(p.ctypes.data, z.ctypes.data, ldz, ndof, n,
                                        dec.ctypes.data if want_decisions else None)
        return (int(h), dec) if want_decisions else int(h)

    def count_streamed_batch(self, pairs, z, n, z_pair_stride=0, threads=0):
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        z = _f32(z)
        ndof, ldz = z.shape
        hits = np.zeros(pairs.size, np.uint64)
        self.lib.orc_count_streamed_batch(pairs.ctypes.data, pairs.size, z.ctypes.data, ldz, z_pair_stride, ndof, n,
                                          hits.ctypes.data, threads)
        return hits

    def mc_thread(self, robot_base, pose, sd, pos, count_in, z, n_batch, n_samples_total, bins, bin_acc):
        robot_base, sd, z = _f32(robot_base), _f32(sd), _f32(z)
        b = _f32(list(bins) + [0.0]); a = _f32(list(bin_acc) + [0.0])
        done = C.c_int(0)
        k = self.lib.orc_mc_thread(robot_base.ctypes.data, float(pose[0]), float(pose[1]), float(pose[2]),
                                   sd.ctypes.data, float(pos[0]), float(pos[1]), int(count_in), z.ctypes.data,
                                   z.shape[1], int(n_batch), int(n_samples_total), b.ctypes.data, a.ctypes.data,
                                   len(bins), C.byref(done))
        return int(k), int(done.value)

    def poly_count_streamed(self, poly_pair, z, want_decisions=False):
        """poly_pair: one element of the 160-byte polygon pair dtype; z: [3, n] float32."""
        p = np.ascontiguousarray(poly_pair).reshape(1)
        assert p.dtype.itemsize == 160
        z = _f32(z)
        n = z.shape[1]
        dec = np.zeros(n, np.uint8) if want_decisions else None
        h = self.lib.orc_poly_count_streamed(p.ctypes.data, z.ctypes.data, n, n, dec.ctypes.data if want_decisions else None)
        return (int(h), dec) if want_decisions else int(h)

    def philox(self, ctr, key):
        ctr = np.ascontiguousarray(ctr, np.uint32); key = np.ascontiguousarray(key, np.uint32)
        out = np.zeros(4, np.uint32)
        self.lib.orc_philox4x32_10(ctr.ctypes.data, key.ctypes.data, out.ctypes.data)
        return out

    def fused_normals(self, seed, pair_id, index, ndof=5):
        z = np.zeros(5, np.float32)
        self.lib.orc_fused_normals(int(seed), int(pair_id), int(index), int(ndof), z.ctypes.data)
        return z[:ndof]

    def count_fused_batch(self, pairs, n_samples, seed, sample_offset=0, pair_id_offset=0, threads=0):
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DTYPE)
        hits = np.zeros(pairs.size, np.uint64)
        self.lib.orc_count_fused_batch(pairs.ctypes.data, pairs.size, int(n_samples), int(seed), int(sample_offset),
                                       int(pair_id_offset), hits.ctypes.data, int(threads))
        return hits


class RefGpu:
    """The unmodified reference compiled for sm_100a (oracle/ref_gpu.cu); needs a GPU."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} missing: run oracle/build_ref.sh where /root/reference exists")
        self.lib = L = C.CDLL(REF_SO)
        vp = C.c_void_p
        L.ref_convex_collide.argtypes = [vp, vp, C.c_int, vp]
        L.ref_rot_trans.argtypes = [vp, vp, vp, vp, C.c_int]
        L.ref_sample_record.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
        L.ref_mc_run.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int,
                                 C.c_int, C.c_int, vp]
        L.ref_write_cp.argtypes = [vp, C.c_int, C.c_int]
        L.ref_dev_sincos.argtypes = [vp, C.c_int, vp, vp, C.c_int]
        L.ref_fast_trig_err.argtypes = [C.c_float, C.c_float, vp, vp]
        L.ref_mc_time.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        L.ref_adaptive_batch.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp]

    @staticmethod
    def _ok(rc):
        if rc != 0:
            raise RuntimeError("reference GPU oracle call failed")

    def convex_collide(self, r1, r2):
        r1, r2 = _f32(r1).reshape(-1, 8), _f32(r2).reshape(-1, 8)
        out = np.zeros(r1.shape[0], np.int32)
        self._ok(self.lib.ref_convex_collide(r1.ctypes.data, r2.ctypes.data, r1.shape[0], out.ctypes.data))
        return out

    def rot_trans(self, r, dx, dy, dt):
        r = _f32(r).reshape(-1, 8).copy()
        dx, dy, dt = _f32(dx), _f32(dy), _f32(dt)
        self._ok(self.lib.ref_rot_trans(r.ctypes.data, dx.ctypes.data, dy.ctypes.data, dt.ctypes.data, r.shape[0]))
        return r

    def sample_record(self, r_in, sd, n_per, seed):
        r_in, sd = _f32(r_in).reshape(-1, 8), _f32(sd).reshape(-1, 5)
        n = r_in.shape[0]
        z = np.zeros((5, n * n_per), np.float32)
        corners = np.zeros((n * n_per, 8), np.float32)
        self._ok(self.lib.ref_sample_record(r_in.ctypes.data, sd.ctypes.data, n, n_per, seed, z.ctypes.data,
                                            corners.ctypes.data))
        return z, corners

    def mc_run(self, robot_base, poses, std_devs, pose_idxs, sd_idxs, positions, cps, bins, bin_acc, n_samples,
               n_batch, seed, record=True):
        robot_base, poses, std_devs = _f32(robot_base), _f32(poses).reshape(-1, 3), _f32(std_devs).reshape(-1, 5)
        pose_idxs, sd_idxs, positions = _f32(pose_idxs), _f32(sd_idxs), _f32(positions).reshape(-1, 2)
        cps = _f32(cps).copy(); bins = _f32(bins); bin_acc = _f32(bin_acc)
        num_left = positions.shape[0]
        done = np.zeros(num_left, np.int32)
        z = np.zeros((5, num_left * n_batch), np.float32) if record else None
        self._ok(self.lib.ref_mc_run(robot_base.ctypes.data, poses.ctypes.data, poses.shape[0], std_devs.ctypes.data,
                                     std_devs.shape[0], pose_idxs.ctypes.data, sd_idxs.ctypes.data,
                                     positions.ctypes.data, cps.ctypes.data, bins.ctypes.data, bins.size,
                                     bin_acc.ctypes.data, done.ctypes.data, n_samples, n_batch, num_left, seed,
                                     z.ctypes.data if record else None))
        return cps, done, z

    def dev_sincos(self, x, fast=False):
        x = _f32(x).ravel()
        s = np.zeros_like(x); c = np.zeros_like(x)
        self._ok(self.lib.ref_dev_sincos(x.ctypes.data, x.size, s.ctypes.data, c.ctypes.data, int(fast)))
        return s, c

    def fast_trig_err(self, lo, hi):
        es, ec = C.c_float(0), C.c_float(0)
        self._ok(self.lib.ref_fast_trig_err(float(lo), float(hi), C.byref(es), C.byref(ec)))
        return float(es.value), float(ec.value)

    def write_cp(self, counts, n_samples):
        c = _f32(counts).copy()
        self._ok(self.lib.ref_write_cp(c.ctypes.data, c.size, n_samples))
        return c

    def adaptive_batch(self, robot_base, poses, std_devs, pose_idxs, sd_idxs, positions, bins, bin_acc, max_samples, seed):
        """The reference's whole adaptive loop (its kernel + thrust compaction). Returns (cp in input order, ms, samples drawn)."""
        robot_base, poses, std_devs = _f32(robot_base), _f32(poses).reshape(-1, 3), _f32(std_devs).reshape(-1, 5)
        pose_idxs, sd_idxs, positions = _f32(pose_idxs), _f32(sd_idxs), _f32(positions).reshape(-1, 2)
        bins, bin_acc = _f32(bins), _f32(bin_acc)
        n = positions.shape[0]
        cp = np.zeros(n, np.float32); ms = C.c_float(0); drawn = C.c_longlong(0)
        self._ok(self.lib.ref_adaptive_batch(robot_base.ctypes.data, poses.ctypes.data, poses.shape[0], std_devs.ctypes.data,
                                             std_devs.shape[0], pose_idxs.ctypes.data, sd_idxs.ctypes.data, positions.ctypes.data, n,
                                             bins.ctypes.data, bins.size, bin_acc.ctypes.data, int(max_samples), int(seed),
                                             cp.ctypes.data, C.byref(ms), C.byref(drawn)))
        return cp, float(ms.value), int(drawn.value)

    def mc_time(self, robot_base, poses, std_devs, pose_idxs, sd_idxs, positions, n_batch, launches, seed):
        robot_base, poses, std_devs = _f32(robot_base), _f32(poses).reshape(-1, 3), _f32(std_devs).reshape(-1, 5)
        pose_idxs, sd_idxs, positions = _f32(pose_idxs), _f32(sd_idxs), _f32(positions).reshape(-1, 2)
        num_left = positions.shape[0]
        ms = C.c_float(0)
        cps = np.zeros(num_left, np.float32)
        self._ok(self.lib.ref_mc_time(robot_base.ctypes.data, poses.ctypes.data, poses.shape[0], std_devs.ctypes.data,
                                      std_devs.shape[0], pose_idxs.ctypes.data, sd_idxs.ctypes.data,
                                      positions.ctypes.data, num_left, n_batch, launches, seed, C.byref(ms),
                                      cps.ctypes.data))
        return float(ms.value), cps
