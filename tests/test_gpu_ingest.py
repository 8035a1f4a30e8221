"""GPU parity: (N2) the library's own ingestion primitives — stable LSD radix sort, exclusive scan (through lgc_csr_build's
chunk table), deduplicated per-user CSR — bit-exact against torch / NumPy (integer work)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,bits", [(1, 8), (31, 16), (4096, 24), (4097, 40), (100_003, 36), (3_000_000, 64), (5000, 1)])
def test_radix_sort_matches_torch_sort(dev, n, bits):
    from lgcnhs_b200 import ops

    g = torch.Generator().manual_seed(n)
    hi = (1 << min(bits, 62)) - 1
    keys = torch.randint(0, hi + 1, (n,), generator=g, dtype=torch.int64)
    if n > 10:
        keys[: n // 3] = keys[n // 3: 2 * (n // 3)]              # many duplicates
    got = ops.sort_u64(keys.clone().to(dev), bits=bits).cpu()
    assert torch.equal(got, torch.sort(keys).values)


def test_radix_sort_is_stable_through_payload(dev):
    """row_order's use: key = (max_deg - deg) << 32 | row — equal degrees keep ascending row order."""
    from lgcnhs_b200 import ops

    g = torch.Generator().manual_seed(0)
    deg = torch.randint(0, 50, (70_000,), generator=g, dtype=torch.int64)
    keys = ((49 - deg) << 32) | torch.arange(deg.numel())
    got = ops.sort_u64(keys.clone().to(dev), bits=32 + 6).cpu() & 0xFFFFFFFF
    ref = torch.argsort(deg, descending=True, stable=True)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("U,M,n", [(7, 5, 0), (50, 80, 1), (300, 500, 12_000), (943, 1682, 90_000), (6040, 3706, 1_100_000)])
def test_seen_csr_matches_numpy(dev, U, M, n):
    from lgcnhs_b200 import ops

    g = np.random.default_rng(U + n)
    u = g.integers(0, U, n)
    i = g.integers(0, M, n)
    if n > 100:
        u[: n // 4], i[: n // 4] = u[n // 4: 2 * (n // 4)], i[n // 4: 2 * (n // 4)]      # duplicate pairs
        u[u == 3] = 4                                                                    # an empty row
    ptr, idx = ops.seen_csr(torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev), U, M)
    key = np.unique(u.astype(np.int64) * M + i)
    ref_ptr = np.r_[0, np.cumsum(np.bincount(key // M, minlength=U))]
    assert np.array_equal(ptr.cpu().numpy(), ref_ptr.astype(np.int32))
    assert np.array_equal(idx.cpu().numpy(), (key % M).astype(np.int32))
    from lgcnhs_b200._lib import LgcnhsError
    if n:
        with pytest.raises(LgcnhsError):
            ops.seen_csr(torch.tensor([U]).to(dev), torch.tensor([0]).to(dev), U, M)      # out-of-range pair is reported


def test_row_order_longest_first(dev):
    from lgcnhs_b200.ops import NormGraph
    from lgcnhs_b200.synth import bipartite_adj, synth_shape

    d = synth_shape("ml-100k")
    adj = torch.from_numpy(bipartite_adj(d.n_users, d.users, d.items)).to(dev)
    n = d.n_users + d.n_items
    gph = NormGraph(adj, n)
    deg = (gph.rowptr[1:] - gph.rowptr[:-1]).to(torch.int64)
    for a, b in ((0, n), (100, 2000)):
        o = gph.row_order(a, b).cpu().to(torch.int64)
        ref = torch.argsort(deg[a:b].cpu(), descending=True, stable=True) + a
        assert torch.equal(o, ref)


def test_unique_and_graph_converters_on_device(dev):
    """lgc_unique_u64 vs torch.unique, and the drop-in utils/graph.py converters on CUDA tensors (own sort/compaction) vs
    the same converters on CPU tensors (torch.unique) — which the CPU tests pin to the reference's outputs."""
    import _stub_const
    from lgcnhs_b200 import ops

    g = torch.Generator().manual_seed(3)
    keys = torch.randint(0, 5000, (200_000,), generator=g, dtype=torch.int64)
    assert torch.equal(ops.unique_u64(keys.clone().to(dev), bits=13).cpu(), torch.unique(keys))
    _stub_const.install()
    from lgcnhs_b200.synth import synth_shape
    from utils import graph

    d = synth_shape("ml-100k")
    ei = torch.from_numpy(np.stack([d.users, d.items]))
    ei = torch.cat([ei, ei[:, :500]], dim=1)                 # duplicates, unsorted
    adj_cpu = graph.convertEdgeIndexToAdjMatrix(d.n_users, d.n_items, ei)
    adj_dev = graph.convertEdgeIndexToAdjMatrix(d.n_users, d.n_items, ei.to(dev))
    assert adj_dev.is_cuda and torch.equal(adj_dev.cpu(), adj_cpu)
    back_cpu = graph.convertAdjMatrixToEdgeIndex(d.n_users, d.n_items, adj_cpu)
    back_dev = graph.convertAdjMatrixToEdgeIndex(d.n_users, d.n_items, adj_dev)
    assert torch.equal(back_dev.cpu(), back_cpu)


def test_transposed_graph_matches_argsort_construction(dev):
    """NormGraph.transposed() (own radix sort of the (source, target) keys, values recomputed as dinv x dinv) against the
    plain construction: a stable argsort of the same keys and the PERMUTED values of the forward graph — integer arrays
    and fp32 values bit-equal, on a non-symmetric graph."""
    from lgcnhs_b200.ops import NormGraph

    gen = torch.Generator().manual_seed(11)
    n, e = 3000, 40000
    src = torch.randint(0, n, (e,), generator=gen)
    dst = (torch.rand(e, generator=gen) ** 2 * n).long().clamp(max=n - 1)      # skewed targets -> some long rows
    key = torch.unique(src * n + dst)
    ei = torch.stack([key // n, key % n]).to(dev)
    g = NormGraph(ei, n)
    t = g.transposed()
    nnz = g.nnz
    deg = (g.rowptr[1:] - g.rowptr[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(n, device=dev), deg)
    cols = g.colidx[:nnz].to(torch.int64)
    order = torch.argsort(cols * n + rows, stable=True)
    assert torch.equal(t.colidx[:nnz], rows[order].to(torch.int32))
    assert torch.equal(t.val[:nnz], g.val[:nnz][order])
    want_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    want_ptr[1:] = torch.cumsum(torch.bincount(cols, minlength=n), 0)
    assert torch.equal(t.rowptr.to(torch.int64), want_ptr)
