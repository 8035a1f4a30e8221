"""torchrun --nproc-per-node N tools/mcast_store_probe.py [shape] : what does the all-gather half of a partitioned
propagation layer cost by itself, and how should a finished row leave the SM?

Every rank owns the rows RowPartitionedPropagation gives it (users and items split separately, longest first) and
  (1) only COPIES them (lgc_probe_row_store, no gather) into every replica, with four store shapes, through the NVSwitch
      multicast address and, for comparison, through the unicast CUDA-IPC mapping of ONE peer;
  (2) runs the real mixed SpMM launch with local stores only / with multicast stores / + the device barrier.
Device time per layer (CUDA events, ranks released together by a device barrier), every rank printed; the ingress rate is
what one GPU RECEIVES per second (all ranks' rows = the whole (n, 64) fp32 table per layer)."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from lgcnhs_b200._lib import check, lib  # noqa: E402
from lgcnhs_b200.dist import PeerGroup, RowPartitionedPropagation, init_dist  # noqa: E402

rank, world, local = init_dist()
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-20m"
d = bench.load_shape(shape, rank, lambda: dist.barrier())
adj_np, _ = bench.train_adj(d)
n = d.n_users + d.n_items
adj = torch.from_numpy(adj_np).to(dev)
torch.manual_seed(42)
x0 = (torch.randn(n, 64) * 0.1).to(dev)
prop = RowPartitionedPropagation(adj, n, 64, mode="p2p", split=d.n_users)
g = prop.g
ranges = [(a, b) for a, b, _ in prop.my_parts]
rows, _lr, _ch = g.row_list(ranges)
n_rows = int(rows.numel())
rows_sorted = torch.sort(rows).values.contiguous()
stream = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    prop.peer_barrier()
    torch.cuda.synchronize()
    dist.barrier()
    prop.peer_barrier()
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]) / reps * 1e3], device=dev)
    every = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(every, t)
    timed.per_rank = [round(float(x.item()), 1) for x in every]
    return max(timed.per_rank)


def store(dst_ptr, mode, row_t):
    check(lib().lgc_probe_row_store(x0.data_ptr(), C.c_void_p(dst_ptr), row_t.data_ptr() if row_t is not None else None,
                                    n_rows, mode, stream()), "probe row store")


out_local = torch.empty_like(x0)
msgs = []
total_mb = n * 256 / 1e6
targets = {"local": out_local.data_ptr()}
if prop.mcast is not None:
    targets["multicast"] = prop.peer_ptrs[0][0]
group = PeerGroup(dev)
uni = torch.empty_like(x0)
uni_ptrs = group.share(uni)
targets["one peer (unicast)"] = uni_ptrs[(rank + 1) % world]
names = {0: "16 lanes x 16 B", 1: "32 lanes x 16 B", 2: "TMA 256 B / row", 3: "TMA 8 KB / 32 rows (sorted rows)"}
quick = os.environ.get("PROBE_QUICK", "0") == "1"
for tname, ptr in (targets.items() if os.environ.get("PROBE_HOL", "0") != "1" else []):
    for mode in ((0, 2) if quick else (0, 1, 2, 3)):
        for order, row_t in (("longest-first", rows), ("sorted", rows_sorted)):
            if (mode == 3 and order != "sorted") or (quick and order == "sorted"):
                continue
            if mode == 3:
                # contiguous tiles: the probe copies rows [0, n_rows) — same bytes, different rows; offset the tables instead
                a0 = ranges[0][0]
                fn = lambda: check(lib().lgc_probe_row_store(x0.data_ptr() + a0 * 256, C.c_void_p(ptr + a0 * 256), None,  # noqa: E731
                                                             ranges[0][1] - a0, 3, stream()), "probe row store")
                rows_here = ranges[0][1] - a0
            else:
                fn = lambda: store(ptr, mode, row_t)  # noqa: E731
                rows_here = n_rows
            t = timed(fn)
            sent = rows_here * 256 / 1e6
            recv = sent * world if tname == "multicast" else sent
            msgs.append(f"  store only -> {tname:20s} {names[mode]:34s} {order:14s}: {t:7.1f} us  ({sent:5.1f} MB sent per rank, "
                        f"{recv / t * 1e3:7.1f} GB/s ingress per GPU)")
# correctness of the multicast TMA store: every replica must hold every rank's rows
if prop.mcast is not None and os.environ.get("PROBE_HOL", "0") != "1":
    prop.bufs[0].zero_()
    torch.cuda.synchronize()
    dist.barrier()
    store(prop.peer_ptrs[0][0], 2, rows)
    prop.peer_barrier()
    torch.cuda.synchronize()
    dist.barrier()
    ok = bool(torch.equal(prop.bufs[0], x0))
    msgs.append(f"  multicast TMA row stores: replica == table on this rank: {ok}")


def spmm(ptrs):
    g.spmm_rows_bcast(x0, x0, 1.0, 1.0, ptrs, ranges)


side = torch.cuda.Stream(device=dev)


def spmm_and_sender():
    """local-store SpMM on the main stream while a SEPARATE kernel on a side stream pushes the rows to the replicas"""
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        store(targets.get("multicast", targets["one peer (unicast)"]), 0, rows)
    spmm([out_local.data_ptr()])
    main.wait_stream(side)


def hol_test():
    """Does a SATURATED exchange stream slow the SpMM down only on the SMs it is issued from?  The local-store SpMM is timed
    (events on the main stream, around the SpMM only) while a persistent sender on a side stream keeps the link busy:
    on n dedicated SMs (exclusive), on n shared SMs, or spread over all SMs (the ordinary grid, back to back)."""
    dst = targets.get("multicast", targets["one peer (unicast)"])
    main = torch.cuda.current_stream()
    out = []

    def sender(n_ctas, excl, passes):
        check(lib().lgc_probe_row_store_persistent(x0.data_ptr(), C.c_void_p(dst), rows.data_ptr(), n_rows, n_ctas, passes,
                                                   excl, stream()), "persistent sender")

    # how fast is the sender alone?
    for n_ctas in (2, 4, 8, 16):
        for excl in (1, 0):
            t = timed(lambda: sender(n_ctas, excl, 4), reps=5) / 4
            sent = n_rows * 256 / 1e6
            out.append(f"  sender alone, {n_ctas:2d} CTAs x 1024 thr {'exclusive SMs' if excl else 'shared SMs   '}: {t:7.1f} us per pass "
                       f"({sent / t * 1e3:6.1f} GB/s egress per GPU, {sent * (world if 'multicast' in targets else 1) / t * 1e3:6.1f} GB/s ingress)")
    t_alone = timed(lambda: spmm([out_local.data_ptr()]))
    out.append(f"  local-store SpMM alone: {t_alone:.1f} us")
    passes_for = max(2, int(t_alone * 2.0 / 60.0) + 1)     # keep the link busy for ~2x the SpMM's duration

    def both(kind, n_ctas):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        reps = 5
        for it in range(reps + 2):
            prop.peer_barrier()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if kind == "all":
                    for _ in range(passes_for):
                        store(dst, 0, rows)
                else:
                    sender(n_ctas, 1 if kind == "excl" else 0, passes_for)
            # give the sender a head start so that it is resident and the link is saturated when the SpMM starts
            torch.cuda._sleep(20000)
            e0.record()
            spmm([out_local.data_ptr()])
            e1.record()
            main.wait_stream(side)
            torch.cuda.synchronize()
            if it >= 2:
                tot += e0.elapsed_time(e1) * 1e3
        t = torch.tensor([tot / reps], device=dev)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        return [round(float(x.item()), 1) for x in every]

    for kind, n_ctas in (("excl", 4), ("excl", 8), ("excl", 16), ("shared", 8), ("all", 0)):
        v = both(kind, n_ctas)
        label = {"excl": f"{n_ctas} dedicated SMs", "shared": f"{n_ctas} CTAs on shared SMs", "all": "ordinary grid on all SMs"}[kind]
        out.append(f"  local-store SpMM while a sender saturates the link from {label:26s}: {max(v):7.1f} us (per rank {v})")
    return out


if os.environ.get("PROBE_HOL", "0") == "1":
    res = hol_test()
    if rank == 0:
        print(f"# {shape}: {world} ranks, multicast={prop.mcast is not None}; head-of-line test")
        print("\n".join(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)

per = {}
t_local = timed(lambda: spmm([out_local.data_ptr()])); per["local"] = timed.per_rank
t_users = timed(lambda: g.spmm_rows_bcast(x0, x0, 1.0, 1.0, [out_local.data_ptr()], ranges[:1])); per["users only"] = timed.per_rank
t_items = timed(lambda: g.spmm_rows_bcast(x0, x0, 1.0, 1.0, [out_local.data_ptr()], ranges[1:])); per["items only"] = timed.per_rank
t_mc = timed(lambda: spmm(prop.peer_ptrs[0])); per["exchange stores"] = timed.per_rank
t_side = timed(spmm_and_sender); per["local + concurrent sender kernel"] = timed.per_rank
t_bar = timed(lambda: (spmm(prop.peer_ptrs[0]), prop.peer_barrier())); per["+barrier"] = timed.per_rank
t_baronly = timed(lambda: prop.peer_barrier(), reps=50)
t_full = timed(lambda: prop.propagate_mean(x0, 3), reps=5) / 3
nnz = sum(int(g.rowptr[b]) - int(g.rowptr[a]) for a, b, _ in prop.my_parts)
if rank == 0:
    print(f"# {shape}: n={n} ({total_mb:.1f} MB per layer), {world} ranks, multicast={prop.mcast is not None}; times = max over ranks")
    print("\n".join(msgs))
    print(f"  SpMM own rows ({nnz / 1e6:.2f} M nnz on rank 0): local stores {t_local:.1f} us | exchange stores {t_mc:.1f} us | "
          f"+ barrier {t_bar:.1f} us | barrier alone {t_baronly:.1f} us | full layer (K=3 call / 3) {t_full:.1f} us", flush=True)
    print(f"  users rows only {t_users:.1f} us, item rows only {t_items:.1f} us, local-store SpMM + concurrent sender kernel "
          f"(side stream) {t_side:.1f} us")
    for k, v in per.items():
        print(f"  per rank, {k}: {v}")
dist.barrier()
dist.destroy_process_group()
