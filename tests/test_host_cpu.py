"""Host-side logic that needs no GPU: .npy I/O and the boost-compatible CLI parser of the C++ programs."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "convex-2d-gpu-collision-detection_b200", "host")
SELFTEST = os.path.join(HOST, "host_selftest")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-s", "-C", HOST, "host_selftest"])


def run(*args):
    return subprocess.run([SELFTEST, *args], capture_output=True, text=True)


@pytest.mark.parametrize("shape", [(7, 5), (1, 4), (100, 3), (4, 1), (0, 5)])
def test_npy_written_by_cpp_is_read_by_numpy(tmp_path, shape):
    f = str(tmp_path / "a.npy")
    assert run("npy-write", f, str(shape[0]), str(shape[1])).returncode == 0
    a = np.load(f)
    want = np.arange(shape[0], dtype=np.float32)[:, None] + np.arange(shape[1], dtype=np.float32)[None, :] / 16
    if shape[1] == 1:
        want = want[:, 0]
    assert a.dtype == np.dtype("<f4") and not np.isfortran(a)
    np.testing.assert_array_equal(a, want)
    assert (os.path.getsize(f) - a.nbytes) % 64 == 0                # header padded like numpy's own writer


@pytest.mark.parametrize("shape", [(12, 5), (3,), (33, 4)])
def test_npy_written_by_numpy_is_read_by_cpp(tmp_path, shape):
    f = str(tmp_path / "b.npy")
    a = np.random.default_rng(0).standard_normal(shape).astype(np.float32)
    np.save(f, a)
    r = run("npy-read", f)
    assert r.returncode == 0, r.stderr
    tok = r.stdout.split()
    assert int(tok[1]) == a.ndim and [int(x) for x in tok[3:3 + a.ndim]] == list(a.shape)
    assert float(tok[tok.index("sum") + 1]) == pytest.approx(float(a.astype(np.float64).sum()), abs=1e-3)


def test_npy_rejects_other_dtypes(tmp_path):
    f = str(tmp_path / "c.npy")
    np.save(f, np.zeros((3, 4), np.float64))
    r = run("npy-read", f)
    assert r.returncode == 2 and "float32" in r.stderr


def test_cli_long_short_equals_and_lists():
    r = run("cli", "--data_dir", "/tmp/x", "-n", "7", "--batch_size=123", "-s", "2", "--max_variance", ".3", ".2", ".1", "0", "0",
            "--shape_variance", "-w", "4.5", "-h", "1.25", "--min_pose", "0.1", "0.1", "-0.5", "--shuffle", "off")
    assert r.returncode == 0, r.stderr
    kv = dict(line.split("=", 1) for line in r.stdout.strip().splitlines())
    assert kv == {"data_dir": "/tmp/x", "num_batches": "7", "batch_size": "123", "start_batch_count": "2",
                  "max_variance": "0.3,0.2,0.1,0,0,", "min_pose": "0.1,0.1,-0.5,", "shape_variance": "1",
                  "robot_width": "4.5", "robot_height": "1.25", "shuffle": "0"}


def test_cli_prefix_unknown_and_help():
    r = run("cli", "--num_b", "3", "--batch", "9")                 # unambiguous prefixes, as boost allows
    assert r.returncode == 0 and "num_batches=3" in r.stdout and "batch_size=9" in r.stdout
    assert run("cli", "--nope", "1").returncode == 2               # unknown option: boost throws, upstream aborts
    assert run("cli", "--s", "1").returncode == 2                  # ambiguous prefix (shape_variance/start_batch_count/shuffle)
    assert run("cli", "--num_batches").returncode == 2             # missing value
    h = run("cli", "--help")
    assert h.returncode == 1 and "-h [ --robot_height ] arg" in h.stdout      # -h is robot_height; --help exits 1
