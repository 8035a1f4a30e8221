"""GPU checks of the device negative sampler (N3, lgc_negative_sample) against the contract of PyG's
structured_negative_sampling as the reference uses it (model/LightGCN/loss.py:46-70, evaluation.py:72):
negatives are never positives of the same user, lie in [0, num_nodes), avoid neg == u when self loops are
excluded, and are uniform over the admissible nodes.  (The random stream itself cannot match torch's.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sampler_contract_and_reproducibility(dev):
    from lgcnhs_b200.sampling import structured_negative_sampling
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape("ml-100k")
    ei = torch.from_numpy(np.stack([d.users, d.items])).to(dev)
    pos = set(zip(d.users.tolist(), d.items.tolist()))
    num_nodes = int(ei.max()) + 1
    g = torch.Generator().manual_seed(7)
    u, p, n = structured_negative_sampling(ei, generator=g)
    assert torch.equal(u, ei[0]) and torch.equal(p, ei[1])
    assert int(n.min()) >= 0 and int(n.max()) < num_nodes                   # range quirk: [0, max id + 1)
    assert all((a, b) not in pos for a, b in zip(u.tolist(), n.tolist()))
    g = torch.Generator().manual_seed(7)
    _, _, n2 = structured_negative_sampling(ei, generator=g)
    assert torch.equal(n, n2)                                               # same seed, same triplets
    _, _, n3 = structured_negative_sampling(ei, generator=g)
    assert not torch.equal(n, n3)                                           # the generator advances
    u, p, n = structured_negative_sampling(ei, contains_neg_self_loops=False, generator=g)
    assert (u != n).all()
    rows = torch.tensor([5, 5, 0, 99999, 17], device=dev)
    u, p, n = structured_negative_sampling(ei, rows=rows, generator=g)
    assert torch.equal(u, ei[0][rows]) and torch.equal(p, ei[1][rows])
    assert all((a, b) not in pos for a, b in zip(u.tolist(), n.tolist()))


def test_sampler_is_uniform_over_admissible_nodes(dev):
    from lgcnhs_b200.sampling import structured_negative_sampling

    # one user with positives {0..9} out of 40 nodes, sampled 120 000 times
    ei = torch.stack([torch.zeros(10, dtype=torch.long), torch.arange(10)])
    ei = torch.cat([ei, torch.tensor([[1], [39]])], dim=1).to(dev)           # second user pins num_nodes = 40
    rows = torch.zeros(120_000, dtype=torch.long, device=dev)
    _, _, n = structured_negative_sampling(ei, rows=rows, generator=torch.Generator().manual_seed(1))
    cnt = torch.bincount(n.cpu(), minlength=40).double()
    assert cnt[:10].sum() == 0
    expect = 120_000 / 30
    chi2 = float(((cnt[10:] - expect) ** 2 / expect).sum())
    assert chi2 < 70.0        # 29 degrees of freedom: P(chi2 > 70) ~ 3e-5


def test_sample_mini_batch_dropin(dev):
    import _stub_const

    _stub_const.install()
    from model.LightGCN.loss import sampleMiniBatch
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape("small")
    ei = torch.from_numpy(np.stack([d.users, d.items])).to(dev)
    pos = set(zip(d.users.tolist(), d.items.tolist()))
    u, p, n = sampleMiniBatch(1024, ei)
    assert u.shape == p.shape == n.shape == (1024,)
    assert all((a, b) in pos for a, b in zip(u.tolist(), p.tolist()))
    assert all((a, b) not in pos for a, b in zip(u.tolist(), n.tolist()))
