#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled for sm_100a (oracle/_ref).

Run on a B200 box:   gpurun -- 'python tools/make_golden.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz into tests/golden/ and commit them.  Inputs are seeded numpy;
outputs are whatever the reference binary computed (its own curand normals included).
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.binding import RefGpu  # noqa: E402

wl = importlib.import_module("convex-2d-gpu-collision-detection_b200.workloads")


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = RefGpu()
    rng = np.random.default_rng(2024)

    # convex_collide: random pairs + near-touching + degenerate + non-finite
    r1, r2 = wl.cfg1_rect_pairs(3000, seed=1)
    base = np.array([-1, -1, 1, -1, 1, 1, -1, 1], np.float32)
    touch1, touch2 = [], []
    for k in range(400):
        off = np.float32(2.0) + np.float32((k - 200) * 1.2e-7)
        t = base.copy(); t[0::2] += off
        ang = np.float32(rng.uniform(0, 6.28)); c, s = np.cos(ang), np.sin(ang)
        R = np.array([[c, -s], [s, c]], np.float32)
        sh = rng.uniform(-3, 3, 2).astype(np.float32)
        touch1.append(((base.reshape(4, 2) @ R.T) + sh).astype(np.float32).reshape(8))
        touch2.append(((t.reshape(4, 2) @ R.T) + sh).astype(np.float32).reshape(8))
    special1 = [base, base, base, np.zeros(8, np.float32), base]
    special2 = [np.full(8, np.nan, np.float32), base + np.float32(np.inf), base * 0, np.zeros(8, np.float32),
                np.array([3, -1, np.nan, -1, 5, 1, 3, 1], np.float32)]
    R1 = np.concatenate([r1, np.array(touch1), np.array(special1)]).astype(np.float32)
    R2 = np.concatenate([r2, np.array(touch2), np.array(special2)]).astype(np.float32)
    np.savez_compressed(os.path.join(out_dir, "ref_convex_collide.npz"), r1=R1, r2=R2, collide=ref.convex_collide(R1, R2))

    # device sinf/cosf
    x = np.concatenate([rng.normal(0, 3, 4000), rng.uniform(-200, 200, 1000), 10.0 ** rng.uniform(-30, 38, 1500),
                        -10.0 ** rng.uniform(-5, 30, 500), [0.0, -0.0, 105614.99, 105615.0, 105615.01, np.pi, np.pi / 2]]).astype(np.float32)
    s, c = ref.dev_sincos(x)
    np.savez_compressed(os.path.join(out_dir, "ref_device_trig.npz"), x=x, sin=s, cos=c)

    # rot_trans_rectangle (standalone contract)
    n = 2000
    r_in = np.stack([wl.cfg1_rect_pairs(n, seed=5)[0]]).reshape(n, 8)
    dx, dy = rng.normal(0, 3, n).astype(np.float32), rng.normal(0, 3, n).astype(np.float32)
    dt = rng.normal(0, 2, n).astype(np.float32)
    np.savez_compressed(os.path.join(out_dir, "ref_rot_trans.npz"), r_in=r_in, dx=dx, dy=dy, dt=dt,
                        r_out=ref.rot_trans(r_in, dx, dy, dt))

    # sample_rectangle on the reference's own curand normals
    n, n_per = 64, 24
    w, h = rng.uniform(0.1, 5, n), rng.uniform(0.1, 5, n)
    rin = np.stack([-w / 2, -h / 2, w / 2, -h / 2, w / 2, h / 2, -w / 2, h / 2], 1).astype(np.float32)
    sd = np.sqrt(rng.uniform(0, 0.3, (n, 5))).astype(np.float32)
    sd[: n // 2, 3:] = 0
    z, corners = ref.sample_record(rin, sd, n_per, seed=77)
    np.savez_compressed(os.path.join(out_dir, "ref_sample_rectangle.npz"), r_in=rin, sd=sd, n_per=n_per, z=z, corners=corners)

    # the MC kernel itself: counts + done flags + the normals it drew
    pairs = wl.dataset_pairs(96, seed=31, shape_variance=True)
    pairs["sd_w"][:48] = 0; pairs["sd_h"][:48] = 0
    robot_base, poses, sds, pi, si, pos = wl.reference_tables(pairs)
    n_batch, n_samples = 200, 1200
    cps_in = rng.integers(0, 1000, pairs.size).astype(np.float32)
    bins = np.array([0, 0.01, 0.1, 1.0], np.float32); acc = np.array([2e-3, 1e-2, 2.2e-2], np.float32)
    cps_out, done, zrec = ref.mc_run(robot_base, poses, sds, pi, si, pos, cps_in, bins, acc, n_samples, n_batch, seed=5)
    np.savez_compressed(os.path.join(out_dir, "ref_mc_kernel.npz"), robot_base=robot_base, poses=poses, std_devs=sds,
                        pose_idxs=pi, sd_idxs=si, positions=pos, cps_in=cps_in, cps_out=cps_out, done=done, z=zrec,
                        n_batch=n_batch, n_samples=n_samples, bins=bins, bin_acc=acc)
    print('mc kernel done flags set:', int(done.sum()), 'of', done.size)
    for f in sorted(os.listdir(out_dir)):
        print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
