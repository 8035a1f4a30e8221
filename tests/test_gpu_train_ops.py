"""GPU parity: (P4/P6/P7) fused BPR forward/backward and Adam vs. torch autograd on the CPU oracle."""
import pytest
import torch

from oracle import lightgcn_oracle as O
from test_gpu_propagation import assert_close

pytestmark = pytest.mark.gpu


def _tables(U, M, D, seed=42):
    g = torch.Generator().manual_seed(seed)
    E = torch.randn(U + M, D, generator=g) * 0.3
    X0 = torch.randn(U + M, D, generator=g) * 0.1
    return E, X0


@pytest.mark.parametrize("U,M,D,B", [(50, 80, 64, 37), (943, 1682, 64, 1024), (300, 200, 32, 256), (64, 64, 128, 100),
                                     (943, 1682, 64, 20000)])
def test_bpr_fwd_bwd(dev, U, M, D, B):
    from lgcnhs_b200 import ops

    E, X0 = _tables(U, M, D)
    g = torch.Generator().manual_seed(1)
    users = torch.randint(U, (B,), generator=g)
    pos = torch.randint(M, (B,), generator=g)
    neg = torch.randint(M, (B,), generator=g)
    eps = 1e-6
    Er, X0r = E.clone().requires_grad_(), X0.clone().requires_grad_()
    loss = O.bpr_loss(Er[users], X0r[users], Er[U + pos], X0r[U + pos], Er[U + neg], X0r[U + neg], eps)
    loss.backward()
    gE = torch.zeros_like(E, device=dev)
    gX0 = torch.zeros_like(X0, device=dev)
    out = ops.bpr_fwd_bwd(E.to(dev), X0.to(dev), U, M, users.to(dev), pos.to(dev), neg.to(dev), eps, gE, gX0)
    assert abs(out[0].item() - loss.item()) <= 1e-5 * abs(loss.item()) + 1e-7
    # every table row sums one term per triplet that touches it, in an order the atomics do not fix: bound the
    # round-off by the sum of the term magnitudes (|term| <= 2 max|E| / B for dL/dE, 2 eps max|X0| for dL/dX0)
    cnt = (torch.bincount(users, minlength=U + M) + torch.bincount(U + pos, minlength=U + M)
           + torch.bincount(U + neg, minlength=U + M)).double()[:, None]
    assert_close(gE, Er.grad, "dL/dE", sum_abs=cnt * 2 * float(E.abs().max()) / B)
    assert_close(gX0, X0r.grad, "dL/dX0", sum_abs=cnt * 2 * eps * float(X0.abs().max()))
    # forward only (calValLoss path): same loss, no gradient buffers
    out2 = ops.bpr_fwd_bwd(E.to(dev), X0.to(dev), U, M, users.to(dev), pos.to(dev), neg.to(dev), eps)
    assert out2[0].item() == out[0].item()


def test_bpr_softplus_threshold(dev):
    """softplus(x) = x above torch's threshold 20; gradient saturates to 1."""
    from lgcnhs_b200 import ops

    D = 64
    u = torch.full((4, D), 1.0)
    p = torch.full((4, D), 0.5)
    n = torch.full((4, D), -0.5)   # s+ - s- = 64 > 20
    z = torch.zeros(4, D)
    rows = [t.requires_grad_() for t in (u.clone(), z.clone(), p.clone(), z.clone(), n.clone(), z.clone())]
    loss = O.bpr_loss(*rows, 1e-6)
    loss.backward()
    grads = [torch.zeros(4, D, device=dev) for _ in range(6)]
    out = ops.bpr_rows([t.detach().to(dev) for t in rows], 1e-6, grads)
    assert abs(out[0].item() - loss.item()) < 1e-5 * abs(loss.item())
    for gk, r in zip(grads, rows):
        assert_close(gk, r.grad, "row grads")


def test_adam_matches_torch(dev):
    from lgcnhs_b200 import ops

    torch.manual_seed(0)
    p = torch.randn(1000, 64) * 0.1
    pr = p.clone().requires_grad_()
    opt = torch.optim.Adam([pr], lr=1e-3)
    pd = p.to(dev)
    m = torch.zeros_like(pd)
    v = torch.zeros_like(pd)
    for step in range(1, 6):
        g = torch.randn(1000, 64) * (0.01 if step % 2 else 1.0)
        pr.grad = g.clone()
        opt.step()
        ops.adam_step(pd, g.to(dev), m, v, step, lr=1e-3)
        assert_close(pd, pr.detach(), f"adam step {step}")
    # odd length exercises the scalar tail
    q = torch.randn(1027)
    qr = q.clone().requires_grad_()
    opt = torch.optim.Adam([qr], lr=0.05)
    qd, m, v = q.to(dev), torch.zeros(1027, device=dev), torch.zeros(1027, device=dev)
    g = torch.randn(1027)
    qr.grad = g.clone()
    opt.step()
    ops.adam_step(qd, g.to(dev), m, v, 1, lr=0.05)
    assert_close(qd, qr.detach(), "adam tail")


@pytest.mark.parametrize("U,M,D,B", [(50, 80, 64, 37), (943, 1682, 64, 1024), (300, 200, 32, 256), (64, 64, 128, 100),
                                     (20, 10, 64, 4000)])
def test_bpr_deterministic_scatter(dev, U, M, D, B):
    """lgc_bpr_fwd_bwd_det: atomic-free gradient scatter — same gradients as autograd on the oracle (duplicates summed in
    entry order), identical to the bit on a second run, and equal to the atomic kernel up to summation order."""
    from lgcnhs_b200 import ops

    E, X0 = _tables(U, M, D)
    g = torch.Generator().manual_seed(2)
    users = torch.randint(U, (B,), generator=g)
    pos = torch.randint(M, (B,), generator=g)
    neg = torch.randint(M, (B,), generator=g)
    eps = 1e-4
    Er, X0r = E.clone().requires_grad_(), X0.clone().requires_grad_()
    loss = O.bpr_loss(Er[users], X0r[users], Er[U + pos], X0r[U + pos], Er[U + neg], X0r[U + neg], eps)
    loss.backward()
    runs = []
    for _ in range(2):
        gE = torch.zeros_like(E, device=dev)
        gX0 = torch.zeros_like(X0, device=dev)
        out = ops.bpr_fwd_bwd_det(E.to(dev), X0.to(dev), U, M, users.to(dev), pos.to(dev), neg.to(dev), eps, gE, gX0)
        runs.append((out.clone(), gE, gX0))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][2], runs[1][2])
    out, gE, gX0 = runs[0]
    # loss = bpr + reg can cancel (B = 4000 triplets over 30 rows: reg ~ +0.77, bpr ~ -0.80): compare the two terms
    bpr_ref = -torch.nn.functional.softplus((E[users] * E[U + pos]).sum(-1) - (E[users] * E[U + neg]).sum(-1)).mean().item()
    reg_ref = loss.item() - bpr_ref
    assert abs(out[1].item() - bpr_ref) <= 1e-5 * abs(bpr_ref) + 1e-7
    assert abs((out[0] - out[1]).item() - reg_ref) <= 2e-5 * abs(reg_ref) + 1e-6
    cnt = (torch.bincount(users, minlength=U + M) + torch.bincount(U + pos, minlength=U + M)
           + torch.bincount(U + neg, minlength=U + M)).double()[:, None]
    # duplicates are added SEQUENTIALLY in fp32 (up to ~800 same-sign regulariser terms per row in the last case): the
    # admissible difference between two fp32 summation orders grows with the number of terms
    n_ops = max(8, int(cnt.max()) // 4)
    assert_close(gE, Er.grad, "det dL/dE", sum_abs=cnt * 2 * float(E.abs().max()) / B, n_ops=n_ops)
    assert_close(gX0, X0r.grad, "det dL/dX0", sum_abs=cnt * 2 * eps * float(X0.abs().max()), n_ops=n_ops)
    untouched = cnt[:, 0] == 0
    assert float(gE[untouched.to(dev)].abs().max() if untouched.any() else 0.0) == 0.0


def test_deterministic_training_is_bit_reproducible(dev):
    """Two trainers, same seed, deterministic scatter, graph replay: bit-identical weights after several steps."""
    import _stub_const

    _stub_const.install()
    from lgcnhs_b200.synth import bipartite_adj, synth_shape
    from lgcnhs_b200.trainer import FusedBPRTrainer
    from model.LightGCN.model import LightGCN

    d = synth_shape("small")
    tr, _, _ = d.split()
    adj = torch.from_numpy(bipartite_adj(d.n_users, d.users[tr], d.items[tr])).to(dev)
    finals = []
    for _ in range(2):
        torch.manual_seed(42)
        m = LightGCN(d.n_users, d.n_items, 64, 3).to(dev)
        t = FusedBPRTrainer(m, adj, lr=1e-2, eps_reg=1e-4, deterministic=True)
        g = torch.Generator().manual_seed(5)
        for s in range(6):
            u = torch.randint(d.n_users, (2048,), generator=g).to(dev)     # many duplicate rows per batch
            p = torch.randint(d.n_items, (2048,), generator=g).to(dev)
            n = torch.randint(d.n_items, (2048,), generator=g).to(dev)
            t.step(u, p, n)
        torch.cuda.synchronize()
        finals.append(t.X0.clone())
    assert torch.equal(finals[0], finals[1])


def test_sparse_first_gradient_layer_equals_dense_backward(dev, monkeypatch):
    """The masked first gradient layer (dL/dE is non-zero on the batch rows only) must leave the training trajectory
    unchanged: same weights, bit for bit, as with LGCNHS_DENSE_BACKWARD=1 — eager steps and graph replays, and the mask is
    all-clear again after every step."""
    import _stub_const

    _stub_const.install()
    from lgcnhs_b200.synth import bipartite_adj, synth_shape
    from lgcnhs_b200.trainer import FusedBPRTrainer
    from model.LightGCN.model import LightGCN

    d = synth_shape("ml-100k")
    tr, _, _ = d.split()
    adj = torch.from_numpy(bipartite_adj(d.n_users, d.users[tr], d.items[tr])).to(dev)
    finals, losses = [], []
    for dense in ("0", "1"):
        monkeypatch.setenv("LGCNHS_DENSE_BACKWARD", dense)
        torch.manual_seed(42)
        m = LightGCN(d.n_users, d.n_items, 64, 3).to(dev)
        t = FusedBPRTrainer(m, adj, lr=1e-2, eps_reg=1e-4, deterministic=True)
        assert t.sparse_backward == (dense == "0")
        g = torch.Generator().manual_seed(5)
        ls = []
        for s in range(5):
            u = torch.randint(d.n_users, (256,), generator=g).to(dev)
            p = torch.randint(d.n_items, (256,), generator=g).to(dev)
            n = torch.randint(d.n_items, (256,), generator=g).to(dev)
            ls.append(t.step(u, p, n).clone())
            assert int(t.row_mask.abs().sum()) == 0
        torch.cuda.synchronize()
        finals.append(t.X0.clone())
        losses.append(torch.stack(ls))
    assert torch.equal(finals[0], finals[1])
    assert torch.equal(losses[0], losses[1])
