"""The CPU restatement (oracle/) against the committed golden vectors, which are outputs of the compiled REFERENCE
(tests/golden/make_golden.py).  This is what pins the oracle; it runs on CPU."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, compare_outputs

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden  # noqa: E402


def _golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return {k: g[k] for k in g.files if k != "trace_sha256"}


@pytest.mark.parametrize("name", sorted(make_golden.FIXTURES))
def test_oracle_matches_reference_golden(name, oracle_player, work_dir):
    from softglrender_b200 import workloads
    from softglrender_b200.scene.trace import read_outputs
    builder, needs_assets = make_golden.FIXTURES[name]
    if needs_assets and workloads.A.find_assets_dir() is None:
        pytest.skip("assets/ not available")
    trace, _ = make_golden.build_trace(name, work_dir)
    out = os.path.join(work_dir, name + ".oracle.out")
    workloads.run_player(oracle_player, trace, out=out, data_dir=work_dir)
    rep = compare_outputs(_golden(name), read_outputs(out), color_frac=0.9999)
    # the restatement follows the reference binary's arithmetic: on the synthetic KATs it is byte-exact
    if name == "kat_1x":
        assert rep["color"]["exact"] == 1.0


def test_oracle_matches_live_reference(oracle_player, work_dir):
    """Where the compiled reference is present, compare live on a trace that is not among the fixtures."""
    from softglrender_b200 import workloads
    from softglrender_b200.scene import synth
    from softglrender_b200.scene.trace import read_outputs
    if not os.path.exists(workloads.REF_PLAYER_ST):
        pytest.skip("oracle/_ref not built in this environment")
    trace = os.path.join(work_dir, "kat_live.sglt")
    synth.kat_trace(140, 110, msaa=True, reverse_z=True, seed=99).save(trace)
    a, b = os.path.join(work_dir, "kat_live.ref.out"), os.path.join(work_dir, "kat_live.oracle.out")
    workloads.run_player(workloads.REF_PLAYER_ST, trace, out=a, data_dir=work_dir)
    workloads.run_player(oracle_player, trace, out=b, data_dir=work_dir)
    compare_outputs(read_outputs(a), read_outputs(b), color_frac=0.9999)
