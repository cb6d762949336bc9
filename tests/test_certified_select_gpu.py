"""The certified select (csrc/gmz_tree.cuh) may only DECIDE what the exact float64 path would decide.
libgmz_verify.so is the same library built with -DGMZ_VERIFY_FAST: every certified decision is re-derived
by the exact path and disagreements are counted on the device (gmz_select_counters)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_certified_decisions_are_never_contradicted_by_the_exact_path():
    from datou_gomoku_muzero_b200 import _build
    so = _build.SO_VERIFY
    if not os.path.exists(so):
        so = _build.build(verify=True)
    env = dict(os.environ, GMZ_LIB=so)
    r = subprocess.run([sys.executable, os.path.join(HERE, "_certified_select_probe.py")], env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    for c in out["cases"]:
        assert c["visits_equal"] and c["moves_equal"] and c["values_equal"], c
        assert c["contradicted"] == 0, c
    assert out["certified"] > 500_000            # the fast path is the one that ran
    assert out["fallback"] < out["certified"] // 1000


def test_production_library_counts_no_verification():
    """The shipped library does not carry the self-check (it would double the select cost)."""
    import torch
    from datou_gomoku_muzero_b200.engine import SearchEngine
    eng = SearchEngine(16, board_size=9, num_simulations=60)
    g = torch.zeros((16, 81), dtype=torch.float64, device="cuda")
    eng.search_e0(g, 3, 16)
    fb, fast, bad = eng.select_counters()
    assert fast == 0 and bad == 0 and fb >= 0
