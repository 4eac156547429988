"""Statistical parity of the random rollouts with the REFERENCE's own rollouts (SURVEY.md 4 plan (iv)).

`tests/golden/ref_rollouts.json` holds value and length of `pure_mcts.MCTS._evaluate_rollout` (pure_mcts.py:86-108,
with `rollout_policy_fn` :7-10 drawing from the global numpy RNG) run unmodified under np.random.seed from three
positions (tests/golden/gen_ref_golden.py).  The engine's rollouts use their own counter-based streams, so parity is
distributional: the reference's sample must be a plausible draw from the engine's outcome and length distributions
(chi-square, p > 1e-4; the fixture and the engine's streams are fixed, so the verdict is deterministic).
Value 2 in the fixture = the reference crashed on a stalemated position (max() of an empty actions() list): the engine
scores such a rollout 0, so both fall in the "no winner" bin.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_rollouts.json")
CHI2_CRIT = {1: 15.14, 2: 18.42, 3: 21.11, 4: 23.51, 5: 25.74, 6: 27.86, 7: 29.88, 8: 31.83, 9: 33.72}   # p = 1e-4


@pytest.fixture(scope="module")
def ref():
    with open(GOLDEN) as f:
        return json.load(f)


def chi2_against(obs_counts, expected_p):
    """chi-square of observed counts against expected proportions; bins with expectation < 5 are pooled."""
    n = obs_counts.sum()
    exp = expected_p * n
    keep = exp >= 5
    o = np.append(obs_counts[keep], obs_counts[~keep].sum())
    e = np.append(exp[keep], exp[~keep].sum())
    if e[-1] < 1e-9:
        o, e = o[:-1], e[:-1]
    return float(((o - e) ** 2 / e).sum()), len(e) - 1


def check(ref_case, big_values, big_plies, name):
    rv = np.array(ref_case["values"])
    rp = np.array(ref_case["plies"])
    obs = np.array([(rv == 1).sum(), (rv == -1).sum(), ((rv == 0) | (rv == 2)).sum()], dtype=np.float64)
    p = np.array([(big_values == 1).mean(), (big_values == -1).mean(), (big_values == 0).mean()])
    c_out, df_out = chi2_against(obs, p)
    edges = np.unique(np.quantile(big_plies, np.linspace(0, 1, 9)[1:-1]))
    big_hist = np.bincount(np.searchsorted(edges, big_plies, side="right"), minlength=len(edges) + 1) / len(big_plies)
    ref_hist = np.bincount(np.searchsorted(edges, rp, side="right"), minlength=len(edges) + 1).astype(np.float64)
    c_len, df_len = chi2_against(ref_hist, big_hist)
    print("%s: reference n=%d P(+1,-1,none)=%s mean plies %.1f | engine n=%d P=%s mean plies %.1f | chi2 outcome %.2f (df %d) "
          "length %.2f (df %d)" % (name, len(rv), np.round(obs / obs.sum(), 3), rp.mean(), len(big_values), np.round(p, 3),
                                   big_plies.mean(), c_out, df_out, c_len, df_len))
    assert c_out <= CHI2_CRIT[max(df_out, 1)], (name, "outcome", c_out)
    assert c_len <= CHI2_CRIT[max(df_len, 1)], (name, "length", c_len)


def test_oracle_rollouts_follow_the_reference_distribution(ref):
    """CPU: the oracle's restatement of the engine's draw procedure (rejection sampling over the precheck superset)."""
    for name, case in ref.items():
        pos = case["position"]
        n = 6000 if name == "start" else 12000
        vals, plies = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        g = O.OracleGame()
        for i in range(n):
            g.set_position(pos["H"], pos["V"], pos["p1"], pos["p2"], pos["w1"], pos["w2"], pos["cur"])
            vals[i], plies[i] = g.rollout(77, 5000 + i, 1000)
        check(case, vals, plies, "oracle/" + name)


@pytest.mark.gpu
def test_kernel_rollouts_follow_the_reference_distribution(ref):
    """GPU: qz_rollout (wall / stuck / pawn kernels) through the C ABI, 65,536 rollouts per position."""
    import torch
    from alphazero_quoridor_b200.quoridor import pack_state
    from alphazero_quoridor_b200.rollout import rollout
    for name, case in ref.items():
        pos = case["position"]
        st = torch.tensor([pack_state(pos["H"], pos["V"], pos["p1"], pos["p2"], pos["w1"], pos["w2"], pos["cur"])],
                          dtype=torch.int64, device="cuda")
        res, plies, _ = rollout(st, per_state=65536, seed=77, rid_base=1 << 20, limit=1000)
        check(case, res.cpu().numpy().astype(np.int64), plies.cpu().numpy().astype(np.int64), "kernel/" + name)
