"""The roofline constants bench.py reports come from profiles/r2_sass_k_count_hotloop.json.  This test re-derives them from
the library that is actually shipped (cuobjdump on libsatmc.so, no GPU needed) and fails when they -- or ptxas's schedule
of the hot loop, which moves throughput by 2 % at equal instruction counts (profiles/r2_codegen_experiments.log) --
have drifted from the committed artefact: regenerate it with
    python tools/sass_count.py --json profiles/r2_sass_k_count_hotloop.json --dump profiles/r2_sass_k_count_hotloop.txt
and re-measure on the GPU."""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"), reason="no cuobjdump")
def test_committed_sass_artifact_matches_the_shipped_library(tmp_path):
    out = tmp_path / "now.json"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "sass_count.py"), "--json", str(out)], stdout=subprocess.DEVNULL)
    now = json.load(open(out))["kernels"]
    ref = json.load(open(os.path.join(ROOT, "profiles", "r2_sass_k_count_hotloop.json")))["kernels"]
    assert set(now) == set(ref)
    for name in ref:
        a = [(L["instr"], L["imad_wide"], L["fp32"], L["mufu"], L["local_spill"], L["fingerprint"]) for L in ref[name]["loops"][:2]]
        b = [(L["instr"], L["imad_wide"], L["fp32"], L["mufu"], L["local_spill"], L["fingerprint"]) for L in now[name]["loops"][:2]]
        assert a == b, f"hot loops of {name} changed: committed {a}, built {b}"
        assert all(x[4] == 0 for x in b), "local-memory traffic (a spill) inside a hot loop"
    # the constants bench.py will print
    sys.path.insert(0, ROOT)
    import bench
    sc = bench.sass_constants()
    assert sc["source"].startswith("profiles/") and 45 <= sc["imad_wide_per_group"] <= 60 and 300 <= sc["instr_per_group"] <= 360


def tile_loop(kernel):
    """the bulk-tensor tile loop of one instantiation: the largest loop that issues UTMALDG and never touches global memory"""
    loops = [L for L in kernel["loops"] if L["opcodes"].get("UTMALDG") and not L["opcodes"].get("LDG")]
    return max(loops, key=lambda L: L["instr"]) if loops else None


@pytest.mark.skipif(shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"), reason="no cuobjdump")
def test_streamed_kernels_are_tma_staged(tmp_path):
    """Every instantiation of the streamed bulk-tensor kernel has a tile loop that requests its data with UTMALDG (TMA),
    waits on an mbarrier (SYNCS), reads shared memory only (no LDG) and keeps nothing in local memory; the committed
    artefact profiles/r2_sass_streamed_tma.json (DESIGN.md's instructions-per-test figures) matches the shipped library."""
    out = tmp_path / "tma.json"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "sass_count.py"), "--kernel", "k_count_streamed_tma",
                           "--min-instr", "40", "--json", str(out)], stdout=subprocess.DEVNULL)
    now = json.load(open(out))["kernels"]
    ref = json.load(open(os.path.join(ROOT, "profiles", "r2_sass_streamed_tma.json")))["kernels"]
    assert set(now) == set(ref) and len(now) == 3                    # <3, 256>, <5, 128> (private banks), <5, 256> (shared bank)
    for name, k in now.items():
        L = tile_loop(k)
        assert L is not None, f"{name}: no tile loop with UTMALDG"
        assert L["opcodes"].get("SYNCS", 0) >= 2 and L["opcodes"].get("LDS", 0) >= 5 and L["local_spill"] == 0, (name, L["opcodes"])
        R = tile_loop(ref[name])
        assert (L["instr"], L["fp32"], L["fp32x2"], L["mufu"]) == (R["instr"], R["fp32"], R["fp32x2"], R["mufu"]), \
            f"tile loop of {name} changed: regenerate profiles/r2_sass_streamed_tma.* and re-measure"
