"""build.log (written by __graft_entry__.build()) ties the shared objects in the tree to the sources they were built from."""
import hashlib
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def test_shared_objects_in_the_tree_are_the_ones_build_log_describes():
    log = os.path.join(ROOT, "build.log")
    if not os.path.exists(log):
        pytest.skip("no build.log: __graft_entry__.build() has not run in this tree")
    rec = json.load(open(log))
    assert rec["arch"].endswith("sm_100a") and rec["make"] and rec["make"][0]["recipe"]
    assert any("compute_100a" in line for line in rec["make"][0]["recipe"])
    stale = [p for p, h in rec["sources_sha256"].items() if not os.path.exists(os.path.join(ROOT, p)) or sha(os.path.join(ROOT, p)) != h]
    if stale:
        pytest.skip(f"sources changed since the last build() ({stale[:3]} ...): build.log describes an older build")
    product = {p: h for p, h in rec["artifacts_sha256"].items() if p.startswith("massively-parallel-gpu-ode-solver_b200/")}
    assert product, "build.log lists no product library"
    for p, h in product.items():
        assert os.path.exists(os.path.join(ROOT, p)), p
        assert sha(os.path.join(ROOT, p)) == h, f"{p} is not the binary build.log describes"
