"""The oracle's stochastic path (oracle/mcts_ref.RefMCTS fed streams of random draws) reproduces the committed outputs of the
UNMODIFIED reference run with the same seeded draws (per-simulation jitter, entropy noise, pruning, direct-model priors)."""
import pytest

import stochastic_cases as S
from oracle.mcts_ref import RefConfig, RefMCTS


@pytest.mark.parametrize("part", range(4))
def test_oracle_reproduces_reference_stochastic_goldens(part):
    d = S.load()
    assert len(d["cases"]) >= 60
    kinds = set()
    for ci, c in enumerate(d["cases"]):
        if ci % 12 != part * 3:        # a quarter of a third of the cases per part keeps the CPU suite short; the GPU suite runs all
            continue
        kinds.add(c["cfg"])
        jit, nrm = S.streams(c["seed"], max(c["expect"]["jitter_used"], 1), max(c["expect"]["normal_used"], 1))
        m = RefMCTS(RefConfig(dirichlet_frac=0.0, playout_random_frac=0.0, num_simulations=c["sims"], **d["configs"][c["cfg"]]), S.backend_of(c),
                    jitter_value=None, jitter_stream=jit, normal_stream=nrm, direct_model=(c["cfg"] == "direct"))
        vc, pi, v = m.run(S.board_of(c), ply=c["ply"])
        S.check(c, vc, pi, v, m._last_root, prior_rtol=0.0)
        assert m.jit_used == c["expect"]["jitter_used"] and m.nrm_used == c["expect"]["normal_used"]
    assert kinds


def test_oracle_reproduces_virtual_loss_goldens():
    """The oracle's in-flight marking == the unmodified reference _select driven with an inflight_counts dict per mini-batch."""
    d = S.load("mcts_vl_golden.json")
    for ci, c in enumerate(d["cases"]):
        if ci % 7:
            continue
        jit, nrm = S.streams(c["seed"], max(c["expect"]["jitter_used"], 1), max(c["expect"]["normal_used"], 1))
        m = RefMCTS(RefConfig(dirichlet_frac=0.0, playout_random_frac=0.0, num_simulations=c["sims"], **d["configs"][c["cfg"]]), S.backend_of(c),
                    jitter_value=None, jitter_stream=jit, normal_stream=nrm, virtual_loss=True)
        vc, pi, v = m.run(S.board_of(c), ply=c["ply"])
        S.check(c, vc, pi, v, m._last_root, prior_rtol=0.0)
        assert m.distinct_rows == c["expect"]["distinct_rows"]
