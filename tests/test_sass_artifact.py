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
