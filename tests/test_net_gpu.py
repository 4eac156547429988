"""GPU: the policy-value net contract (policy_value_net.py:127-164) and the net-driven search.

The net stays in PyTorch (bf16, channels_last, eval-mode BN for batched inference).  The only numerics claim is
against a plain fp32 PyTorch forward of the SAME weights: |prob diff| <= 2e-2 and |value diff| <= 5e-2 (bf16)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net():
    from alphazero_quoridor_b200.policy_value_net import PolicyValueNet
    torch.manual_seed(0)
    return PolicyValueNet(use_gpu=True)


def test_param_count_and_state_dict_keys(net):
    sd = net.get_policy_param()
    assert sum(p.numel() for p in net.policy_value_net.parameters()) == 453041
    for k in ("conv1.weight", "bn1.running_mean", "res1.conv1.weight", "res5.bn2.bias", "conv2.weight", "fc1.weight",
              "fc2.bias", "conv3.weight", "bn3.weight", "fc3.weight"):
        assert k in sd


def test_batched_bf16_inference_matches_fp32(net):
    from alphazero_quoridor_b200.quoridor import BatchedQuoridor
    from alphazero_quoridor_b200.synthetic import midgame_positions
    states = midgame_positions(2048, seed=3, min_plies=0, max_plies=60)
    probs, value = net.evaluate_states(states)
    assert probs.shape == (2048, 140) and value.shape == (2048,) and probs.dtype == torch.float32
    env = BatchedQuoridor(2048, states=states)
    x = env.encode(dtype=torch.float32)
    ref = net.policy_value_net.eval()
    with torch.no_grad():
        logp, v = ref(x)
    net.policy_value_net.train()
    assert (probs - logp.exp()).abs().max().item() <= 2e-2
    assert (value - v.view(-1)).abs().max().item() <= 5e-2
    np.testing.assert_allclose(probs.sum(1).cpu().numpy(), 1.0, atol=2e-2)


def test_policy_value_fn_contract(net):
    """policy_value_net.py:145-164: (zip(legal, probs[legal]) NOT renormalised, value)."""
    from alphazero_quoridor_b200.quoridor import Quoridor
    g = Quoridor()
    act_probs, value = net.policy_value_fn(g)
    pairs = list(act_probs)
    assert [a for a, _ in pairs] == g.actions()
    mass = sum(float(p) for _, p in pairs)
    assert 0.5 < mass <= 1.0 + 1e-3
    assert -1.0 <= float(value) <= 1.0
    p, v = net.policy_value(np.stack([g.state(), g.state()]))
    assert p.shape == (2, 140) and v.shape == (2, 1)


def test_net_driven_search_batched_and_single(net):
    from alphazero_quoridor_b200 import mcts, tree
    from alphazero_quoridor_b200.quoridor import Quoridor
    from alphazero_quoridor_b200.synthetic import midgame_positions
    n = 256
    states = midgame_positions(n, seed=9, min_plies=0, max_plies=30)
    eng = tree.BatchedMCTS(n, tree.NetEvaluator(net), c_puct=5, n_playout=32, leaves_per_game=4, reuse_tree=True)
    eng.reset(states)
    eng.search()
    visits, probs, rootn = eng.root_stats(temp=1.0)
    assert (rootn == 32).all() and (visits.sum(1) == 31).all()
    moves = eng.choose(mode=2, temp=1.0, seed=1)
    legal = torch.gather(visits, 1, moves.long().clamp(min=0).unsqueeze(1)).squeeze(1)
    assert ((moves >= 0) & (moves < 140)).all()
    eng.advance(moves)
    eng.search()
    assert eng.overflow_count() == 0
    # the reference API on top of the same path
    player = mcts.MCTSPlayer(net.policy_value_fn, c_puct=5, n_playout=16, is_selfplay=1)
    g = Quoridor()
    np.random.seed(1)
    move, pr = player.choose_action(g, temp=1.0, return_prob=1)
    assert move in g.actions() and abs(pr.sum() - 1) < 1e-9


def test_train_step_runs(net):
    """policy_value_net.py:166-192 (next-row item; here only: it runs and returns floats)."""
    from alphazero_quoridor_b200.policy_value_net import PolicyValueNet
    torch.manual_seed(1)
    t = PolicyValueNet(use_gpu=True)
    rng = np.random.RandomState(0)
    s = rng.randint(0, 2, size=(16, 26, 9, 9)).astype(np.float64)
    p = rng.dirichlet(np.ones(140), size=16)
    z = rng.choice([-1.0, 1.0], size=16)
    l0, e0 = t.train_step(s, p, z, 2e-3)
    for _ in range(5):
        l1, e1 = t.train_step(s, p, z, 2e-3)
    assert isinstance(l0, float) and isinstance(e0, float) and l1 < l0
