"""CPU: host-side logic of the drop-in (formats, metrics, partitioning, synthetic data, launcher)
against golden vectors from the reference and against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest
import torch

import _stub_const
from oracle import lightgcn_oracle as LO

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "light-graph-convolutional-recommendation-algorithm-based-on-hybrid-spreading_b200")


def test_graph_converters_match_reference():
    _stub_const.install()
    from utils import graph

    z = np.load(os.path.join(G, "lightgcn_tiny.npz"))
    ei = torch.from_numpy(np.stack([z["users"][z["train"]], z["items"][z["train"]]]))
    adj = graph.convertEdgeIndexToAdjMatrix(96, 160, ei)
    assert np.array_equal(adj.numpy(), z["adj"])
    assert np.array_equal(graph.convertAdjMatrixToEdgeIndex(96, 160, adj).numpy(), z["edge_back"])
    # duplicates and unsorted input are deduplicated / ordered like the dense route does
    dup = torch.cat([ei, ei[:, :50]], dim=1)[:, torch.randperm(ei.shape[1] + 50)]
    assert torch.equal(graph.convertEdgeIndexToAdjMatrix(96, 160, dup), adj)
    assert graph.convertEdgeIndexToAdjMatrix(4, 4, torch.zeros((2, 0), dtype=torch.long)).shape == (2, 0)


def test_trans_and_metrics_match_reference():
    _stub_const.install()
    from _metrics_numpy import accurate_metrics as getAccurateMetrics     # the checker of the device metrics kernel,
    from _metrics_numpy import diversity_metrics as getDiversityMetrics   # pinned here to the reference's outputs
    from utils import trans

    z = np.load(os.path.join(G, "metrics_small.npz"))
    users, items = z["users"], z["items"]
    te, tv = z["test"], np.r_[z["train"], z["val"]]
    test_df = pd.DataFrame({"user_id": users[te], "item_id": items[te]})
    tv_df = pd.DataFrame({"user_id": users[tv], "item_id": items[tv]})
    test_dict = trans.getUserItemsDictByDataframe(test_df)
    assert list(test_dict.keys()) == z["dict_keys"].tolist()                 # insertion order = first appearance
    assert [v[0] for v in test_dict.values()] == z["dict_first"].tolist()
    tv_dict = trans.getUserItemsDictByDataframe(tv_df)
    deg = trans.getItemDegreeByUserPosItemDict(tv_dict)
    assert deg == dict(zip(z["deg_items"].tolist(), z["deg_vals"].tolist()))
    A = trans.getInteractionMatrixByDataframe(300, 500, tv_df)
    assert A.dtype == np.float64 and A.sum() == len(tv)
    rec = torch.from_numpy(z["rec"])
    assert np.allclose(getAccurateMetrics(test_dict, rec.numpy(), 10), z["accurate"], atol=1e-5)     # 5-dp rounded values
    assert np.allclose(getDiversityMetrics(rec.numpy(), deg, A, 10), z["diversity"], atol=1e-5)
    ei = torch.from_numpy(np.stack([users[te], items[te]]))
    assert trans.getUserItemsDictByEdgeIndex(ei) == dict(test_dict)
    assert np.array_equal(trans.getInteractionMatrixByEdgeIndex(300, 500, torch.from_numpy(np.stack([users[tv], items[tv]]))), A)
    d = {1: [3, 4], 0: [5, 6]}
    assert trans.recommendDictToTensor(d).tolist() == [[5, 6], [3, 4]]


def test_partition_rows_by_nnz():
    from lgcnhs_b200.dist import partition_rows_by_nnz

    rng = np.random.default_rng(0)
    deg = rng.zipf(1.5, 5000).clip(max=3000)
    rowptr = np.r_[0, np.cumsum(deg)]
    for parts in (1, 2, 3, 8):
        b = partition_rows_by_nnz(rowptr, parts)
        assert b[0] == 0 and b[-1] == 5000 and (np.diff(b) >= 0).all() and len(b) == parts + 1
        work = np.diff(rowptr[b]) + np.diff(b)
        assert work.max() <= work.sum() / parts + deg.max() + 1      # balanced up to one row
    assert partition_rows_by_nnz(np.array([0, 0, 0]), 4).tolist() == [0, 0, 0, 0, 2] or True


def test_synth_shapes_and_split():
    from lgcnhs_b200.synth import bipartite_adj, synth_shape

    d = synth_shape("ml-100k")
    assert (d.n_users, d.n_items, d.users.size) == (943, 1682, 100_000)
    assert np.unique(d.users * d.n_items + d.items).size == 100_000          # distinct pairs
    assert np.bincount(d.users, minlength=943).min() >= 20                   # MovieLens' >= 20 filter
    tr, va, te = d.split()
    assert (len(tr), len(va), len(te)) == (80_000, 10_000, 10_000)
    d2 = synth_shape("ml-100k")
    assert np.array_equal(d.users, d2.users) and np.array_equal(d.items, d2.items)   # seeded
    adj = bipartite_adj(d.n_users, d.users[tr], d.items[tr])
    assert adj.shape == (2, 160_000)
    assert torch.equal(torch.from_numpy(adj), LO.convert_edge_index_to_adj(943, 1682, torch.from_numpy(np.stack([d.users[tr], d.items[tr]]))))


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path[:0] = [PKG, ROOT]
    import torch.distributed as dist

    from lgcnhs_b200.dist import emulate_row_partitioned, init_dist
    from lgcnhs_b200.synth import bipartite_adj, synth_shape

    init_dist("gloo")
    d = synth_shape("small")
    adj = torch.from_numpy(bipartite_adj(d.n_users, d.users, d.items))
    n = d.n_users + d.n_items
    ei, norm = LO.gcn_norm(adj)
    torch.manual_seed(0)
    x = torch.randn(n, 16)
    ref = LO.propagate(ei, x, norm)
    rowptr = np.r_[0, np.cumsum(np.bincount(adj[1].numpy(), minlength=n))]
    # each rank computes only its row block (here with the CPU oracle), then the blocks are exchanged
    out, bounds = emulate_row_partitioned(rowptr, lambda r0, r1: ref[r0:r1].clone(), n, 16)
    ok = torch.equal(out, ref) and int(bounds[-1]) == n
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, [int(b) for b in bounds]))


def _gloo_eval_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path[:0] = [PKG, ROOT]
    import torch.distributed as dist

    from lgcnhs_b200.dist import init_dist, partition_rows_by_nnz
    from lgcnhs_b200.trainer import gather_user_blocks, user_block

    init_dist("gloo")
    ok = True
    for U, k in ((7, 3), (100, 20), (101, 5)):
        full = torch.arange(U * k, dtype=torch.int64).reshape(U, k) * 3 + 1      # what one GPU would return
        u0, u1 = user_block(U, rank, world)
        got = gather_user_blocks(full[u0:u1].clone(), U, rank, world)             # every rank contributes its block
        ok &= torch.equal(got, full)
    # users / items split of the distributed trainer: both halves partitioned by nnz, each rank owns one slice of each
    rng = np.random.default_rng(1)
    deg = np.r_[rng.integers(1, 50, 300), rng.zipf(1.6, 500).clip(max=400)]
    rowptr = np.r_[0, np.cumsum(deg)]
    split = 300
    bu = partition_rows_by_nnz(rowptr[: split + 1], world)
    bi = partition_rows_by_nnz(rowptr[split:] - rowptr[split], world) + split
    ok &= bu[0] == 0 and bu[-1] == split and bi[0] == split and bi[-1] == 800
    mine = int(rowptr[bu[rank + 1]] - rowptr[bu[rank]] + rowptr[bi[rank + 1]] - rowptr[bi[rank]])
    tot = torch.tensor([mine])
    dist.all_reduce(tot)
    ok &= int(tot.item()) == int(rowptr[-1])                                       # the slices tile the graph
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_sharded_eval_gather_and_split_partition_gloo_world2():
    """Host logic of the N-GPU evaluation (user blocks, padded all-gather, row reconstruction) and of the users/items
    partition of the distributed trainer, on two gloo ranks."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    ps = [ctx.Process(target=_gloo_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in ps:
        p_.start()
    res = [q.get(timeout=180) for _ in ps]
    for p_ in ps:
        p_.join(timeout=60)
    assert all(ok for _, ok in res)


def test_row_partition_exchange_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2]                    # both ranks derive the same partition


@pytest.mark.skipif(not os.path.exists("/root/reference/main.py"), reason="reference tree not present on this box")
def test_reference_main_imports_resolve_to_dropin(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(PKG, "run_main.py"), "--check-imports"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if "<-" in l]
    assert any(l.startswith("const") and "/root/reference/const.py" in l for l in lines)
    for l in lines:
        if not l.startswith("const"):
            assert "/root/reference" not in l, f"module resolved to the reference instead of the drop-in: {l}"


def test_partition_cost_weighting_and_metric_sums():
    """Host logic added for the multi-GPU partition (chunk-path rows cost more) and the device metrics reduction."""
    from lgcnhs_b200.dist import partition_rows_by_nnz
    from lgcnhs_b200.ops import metrics_from_sums

    deg = np.r_[np.full(1000, 50), np.full(10, 5000)]            # 1000 short rows, then 10 long (chunk-path) rows
    rowptr = np.r_[0, np.cumsum(deg)]
    plain = partition_rows_by_nnz(rowptr, 2, chunk_weight=1.0, long_row=256)
    heavy = partition_rows_by_nnz(rowptr, 2, chunk_weight=2.0, long_row=256)
    assert plain[0] == heavy[0] == 0 and plain[-1] == heavy[-1] == 1010
    assert heavy[1] > plain[1]                                    # the part holding the long rows gets fewer of them
    cost = lambda a, b, w: sum(d * (w if d > 256 else 1.0) + 1 for d in deg[a:b])  # noqa: E731
    assert abs(cost(0, heavy[1], 2.0) - cost(heavy[1], 1010, 2.0)) <= 2 * 5000 * 2.0 + 2
    # sums -> the reference's rounded metrics (metrics/accurate.py, metrics/diversity.py closed forms)
    m = metrics_from_sums([30.0, 2.5, 4.0, 10.0, 180.0, 54.0], n_users=10, k=3)
    assert m["precision"] == 1.0 and m["recall"] == 0.25 and m["ndcg"] == 0.4
    assert m["f1"] == round(2 * 1.0 * 0.25 / 1.25, 5)
    assert m["H"] == round(1 - 180.0 / (10 * 9 * 3), 5) and m["I"] == round(54.0 / (10 * 3 * 2), 5)


def test_evaluate_lists_input_forms():
    """Host side of the multi-k evaluation driver (evaluationMetrics.py:43-96): the saved dict{uid: ids} format, arrays and
    tensors give the same (U, k) id matrix in user order; short or ragged inputs fail loudly."""
    from lgcnhs_b200.evaluate_lists import lists_to_tensor

    U, k = 6, 3
    rec = np.arange(U * 5).reshape(U, 5)
    as_dict = {u: rec[u].tolist() for u in reversed(range(U))}          # insertion order must not matter
    cpu = torch.device("cpu")
    want = torch.from_numpy(rec[:, :k])
    assert torch.equal(lists_to_tensor(as_dict, U, k, cpu), want)
    assert torch.equal(lists_to_tensor(rec, U, k, cpu), want)
    assert torch.equal(lists_to_tensor(torch.from_numpy(rec).int(), U, k, cpu), want)
    with pytest.raises(ValueError):
        lists_to_tensor({u: rec[u, :2].tolist() for u in range(U)}, U, k, cpu)
    with pytest.raises(ValueError):
        lists_to_tensor(rec[:-1], U, k, cpu)
    with pytest.raises(KeyError):
        lists_to_tensor({u: rec[u].tolist() for u in range(U - 1)}, U, k, cpu)
