#!/usr/bin/env python
"""bench.py — the driver-facing benchmark of the LGCNHS hot paths on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  (reference arm: the reference's CPU
                                                              algorithm on the box's host cores)

Primary metric (BASELINE.json: "LightGCN prop GB/s (frac of HBM peak)"): one step = one K=3,
D=64 fused propagation + layer mean over the TRAIN graph of the ML-20M shape
(138 493 x 26 744, 20 M interactions -> nnz = 32 M) — BASELINE config 5's propagation, the
largest propagation workload and one that fits a single B200.  value = algorithmic bytes
(264 B per non-zero + 260 B per node, per layer) / device time.
With N GPUs the rows are partitioned by nnz; each rank computes its rows and stores them into
every peer's replica over NVLink inside the SpMM kernel (fused all-gather), one barrier per
layer.  scaling = "strong" (the graph is fixed).

The same JSON line carries the two other quantities the metric names, measured on BASELINE
config 2 (ML-1M shape, hybrid spreading, top-20): "spreading": W GEMM TFLOP/s and users/s.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "light-graph-convolutional-recommendation-algorithm-based-on-hybrid-spreading_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

K_LAYERS, DIM = 3, 64


def ncu_traffic(key: str):
    """roofline.traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel.  It cannot be
    measured in this process (a number taken under a profiler is never a bench value), so it is read from the
    COMMITTED capture summary profiles/ncu_traffic.json — written by tools/ncu_traffic.py from an `ncu --set full`
    CSV kept beside it — and only if that CSV's sha256 still matches; otherwise the field is null."""
    import hashlib

    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            table = json.load(f)
        ent = table[key]
        with open(os.path.join(ROOT, "profiles", ent["source"]), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != ent["sha256"]:
                return None, {"error": "profiles/%s changed since ncu_traffic.json was written" % ent["source"]}
        return ent["dram_bytes_per_launch"], {k: ent[k] for k in ("source", "sha256", "kernel", "launches", "captured_with")
                                              if k in ent}
    except Exception as e:
        return None, {"error": "no committed ncu capture for %s (%s)" % (key, type(e).__name__)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class Clocks:
    """nvidia-smi clock / throttle-reason sampler: one streaming process (-lms 20) started before the
    warm-up; summary() keeps the samples whose timestamp falls inside the timed region (all samples
    taken under load if the region is shorter than the sampling period)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        import tempfile

        self.path = tempfile.mktemp(prefix="lgc_clocks_", suffix=".csv")
        self.f = open(self.path, "w")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        self.t0 = self.t1 = self.p0 = self.p1 = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.f.close()

    def probe(self, fn, sync, seconds: float = 0.6):
        """The timed region of a fast step is shorter than nvidia-smi's sampling period: repeat the identical step
        loop, untimed, for `seconds` so that the sampler sees the same load (used only when the timed region itself
        caught fewer than 3 samples)."""
        self.p0 = time.time()
        while time.time() - self.p0 < seconds:
            for _ in range(10):
                fn()
            sync()
        self.p1 = time.time()

    def summary(self):
        import datetime

        rows = []
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.strip().split(",")]
                if len(c) < 7:
                    continue
                try:
                    ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(c[1]), float(c[2]), c[3:7]))
                except ValueError:
                    continue
        except OSError:
            pass
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        where = "timed region"
        if len(inside) < 3 and getattr(self, "p0", None) is not None:
            inside = [r for r in rows if self.p0 <= r[0] <= self.p1 + 0.02]
            where = "identical untimed repeat of the step loop right after the timed region"
        use = inside if inside else rows[-5:]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in use for n, v in zip(names, r[3]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median([r[1] for r in use])) if use else None,
                "sm_max_mhz": max((r[2] for r in use), default=None), "reasons": reasons,
                "samples": len(inside), "samples_total": len(rows), "sampled_during": where}


# ------------------------------------------------------------------------------------------
# workload construction (synthetic, seeded; cached per box under /tmp)
# ------------------------------------------------------------------------------------------
def load_shape(name: str, rank: int = 0, barrier=None):
    from lgcnhs_b200.synth import Interactions, synth_shape

    path = f"/tmp/lgcnhs_synth_{name}_42.npz"
    if rank == 0 and not os.path.exists(path):
        d = synth_shape(name)
        np.savez(path + ".tmp.npz", users=d.users, items=d.items, n=np.array([d.n_users, d.n_items]))
        os.replace(path + ".tmp.npz", path)
    if barrier is not None:
        barrier()
    z = np.load(path)
    return Interactions(int(z["n"][0]), int(z["n"][1]), z["users"], z["items"])


def train_adj(d):
    from lgcnhs_b200.synth import bipartite_adj

    tr, va, te = d.split()
    return bipartite_adj(d.n_users, d.users[tr], d.items[tr]), (tr, va, te)


def bipartite(d, idx):
    from lgcnhs_b200.synth import bipartite_adj

    return bipartite_adj(d.n_users, d.users[idx], d.items[idx])


def prop_bytes(nnz: int, n: int, layers: int = K_LAYERS, dim: int = DIM) -> int:
    return layers * (nnz * (4 + 4 + 4 * dim) + n * (4 * dim + 4) + 4)


# ------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port; PyG is not installable)
# ------------------------------------------------------------------------------------------
def cpu_prop_sample(adj: np.ndarray, n_users: int, n_items: int, steps: int, warmup: int):
    """One PyG-equivalent propagation layer (gcn_norm + index_select/mul/scatter_add,
    model/LightGCN/model.py:53,62,84) on the full graph, all host threads."""
    from oracle import lightgcn_oracle as LO

    ei = torch.from_numpy(adj)
    torch.manual_seed(42)
    x = torch.empty(n_users + n_items, DIM).normal_(std=0.1)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        e, norm = LO.gcn_norm(ei)           # the reference recomputes the norm on every forward
        LO.propagate(e, x, norm)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    sec = float(np.mean(ts))
    return prop_bytes(adj.shape[1], n_users + n_items, layers=1) / sec / 1e9, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm must use the box's host cores at every N
    # (the other ranks exit at once, so rank 0 has the whole host)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, cores))
    d = load_shape(args.shape)
    adj, _ = train_adj(d)
    steps, warmup = max(1, min(args.steps, 40)), max(0, min(args.warmup, 3))   # ~1.5 s per step on 16 cores
    gbs, sec = cpu_prop_sample(adj, d.n_users, d.n_items, steps, warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "LightGCN prop GB/s", "value": round(gbs, 3), "unit": "GB/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": round(sec * 1e3, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"LightGCN propagation D={DIM}, {args.shape} shape train graph "
                               f"(nnz={adj.shape[1]}, N={d.n_users + d.n_items})"},
        "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": "1 of 3 propagation layers over the full graph (gcn_norm + index_select + "
                                   "scatter_add, PyG-equivalent oracle port), algorithmic bytes of 1 layer / time"},
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def install_cfg(model: str, dataset: str = "movielens", k: int = 20, lam: float = 0.3):
    """The drop-in modules read the reference's `const.cfg` (const.py:111-190).  /root/reference is not on the GPU box,
    so the bench installs a module of that name with the attributes the recommend* entry points read."""
    import tempfile
    import types

    root = tempfile.mkdtemp(prefix="lgc_bench_cfg_")
    paths = {n: os.path.join(root, n) + "/" for n in ("log", "preprocess", "recommend", "model", "evaluation", "pictures")}
    for q in paths.values():
        os.makedirs(q, exist_ok=True)
    cfg = types.SimpleNamespace(
        DATA_SET=dataset, LOG={"file_path": paths["log"]}, PREPROCESSING={"seed": 42, "save_path": paths["preprocess"]},
        RECOMMEND={"k": k, "save_path": paths["recommend"]}, EVALUATION={"save_path": paths["evaluation"]},
        PICTURES={"save_path": paths["pictures"]},
        MODEL={"name": model, "save_path": paths["model"],
               "HyperParameter": {"seed": 42, "embedding_dim": DIM, "layers": K_LAYERS, "lr": 1e-3, "gamma": 0.95, "epochs": 4,
                                  "epoch_per_eval": 2, "epoch_per_lr_decay": 2, "batch_size": 1024, "epsilon": 1e-6,
                                  "lambda": lam}})
    mod = types.ModuleType("const")
    mod.cfg = cfg
    sys.modules["const"] = mod
    for name in [m for m in sys.modules if m.split(".")[0] in ("model", "utils", "metrics", "processing")]:
        del sys.modules[name]
    return cfg


def frames_of(d, *index_sets):
    import pandas as pd

    return [pd.DataFrame({"user_id": d.users[ix], "item_id": d.items[ix]}) for ix in index_sets]


def wall_ms(fn, reps: int, warm: int = 1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def spreading_e2e(d, k: int = 20):
    """e2e of BASELINE config 2 through the reference-facing plugin call: recommendSpreadMethod(U, M, train_df, val_df,
    "HybridS") with HOST DataFrames in and the reference's dict{uid: [k ids]} out (reference recommend.py:59-115).
    Inside the timed region: interaction upload, degree / operand packing, G GEMM, HybridS scaling, F GEMM, filtered
    top-k, id download, dict construction and the np.save side effect."""
    install_cfg("HybridS", k=k, lam=0.3)
    from model.SpreadMethod.recommend import recommendSpreadMethod

    tr, va, _ = d.split()
    train_df, val_df = frames_of(d, tr, va)
    out = {}

    def call():
        out["rec"] = recommendSpreadMethod(d.n_users, d.n_items, train_df, val_df, "HybridS")
    ms = wall_ms(call, reps=3)
    assert len(out["rec"]) == d.n_users
    # the same call without the reference's np.save side effect (a pickle of 120 k np.int64 scalars dominates the wall clock)
    import model.SpreadMethod.recommend as R

    keep = R._save
    R._save = lambda rec: None
    try:
        ms_nosave = wall_ms(call, reps=3)
    finally:
        R._save = keep
    return {"ms": round(ms, 3), "users_per_s": round(d.n_users / (ms * 1e-3), 1), "unit": "users/s",
            "ms_without_np_save": round(ms_nosave, 3), "users_per_s_without_np_save": round(d.n_users / (ms_nosave * 1e-3), 1),
            "h2d_bytes_per_step": int(2 * 8 * (tr.size + va.size)), "d2h_bytes_per_step": int(d.n_users * k * 8),
            "what": "recommendSpreadMethod(U, M, train_df, val_df, 'HybridS'): host DataFrames -> dict{uid: top-20}, wall "
                    "clock incl. upload, packing, G, scale, F, top-k, download, dict + np.save (median of 3)"}


def fusion_leg(dev):
    """BASELINE config 3: SpreadLightGCNOpti on the Douban shape — (layer-0 score of the feature-initialised LightGCNOpti,
    masked to -1024) * (A . HybridS(lambda)) -> filtered top-20, one B200."""
    install_cfg("SpreadLightGCNOpti", dataset="douban", k=20, lam=0.3)
    from lgcnhs_b200 import fusion
    from model.LightGCNOpti.model import LightGCNOpti

    d = load_shape("douban")
    tr, va, _ = d.split()
    train_df, val_df = frames_of(d, tr, va)
    rng = np.random.default_rng(3)
    torch.manual_seed(42)
    model = LightGCNOpti(d.n_users, d.n_items, DIM, K_LAYERS, torch.from_numpy(rng.random((d.n_users, 29)).astype(np.float32)),
                         torch.from_numpy(rng.random((d.n_items, 31)).astype(np.float32))).to(dev)
    out = {}

    def call():
        out["idx"] = fusion.fused_recommend(model, d.n_users, d.n_items, train_df, val_df, 0.3, 20).cpu()
    ms = wall_ms(call, reps=5)
    flops = 2.0 * d.n_items * d.n_items * d.n_users * 2
    return {"workload": f"SpreadLightGCNOpti fused recommend, douban shape ({d.n_users}x{d.n_items}, nnz(A)={tr.size + va.size})",
            "ms": round(ms, 3), "users_per_s": round(d.n_users / (ms * 1e-3), 1),
            "useful_tflops": round(flops / (ms * 1e-3) / 1e12, 2),
            "what": "host DataFrames -> (U, 20) ids on the host: upload, G = A^T K_u^-1 A (16 000^2), HybridS scaling, "
                    "F = A.W, fused layer-0 score x F top-20 (lgc_score_topk mul), download; wall clock, median of 5"}


def spreading_leg(dev, steps: int, warmup: int):
    """BASELINE config 2: ML-1M shape, G once, then per lambda: scale + F = A.W + filtered top-20."""
    from lgcnhs_b200 import ops

    d = load_shape("ml-1m")
    tr, va, _ = d.split()
    sel = np.concatenate([tr, va])
    users = torch.from_numpy(d.users[sel]).to(dev)
    items = torch.from_numpy(d.items[sel]).to(dev)
    eng = ops.SpreadingEngine(d.n_users, d.n_items, users, items)
    U, M = d.n_users, d.n_items
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    # G build: operand packing (memset + scatter of A^T and the digit planes) and the int8 tcgen05 GEMM, timed apart
    ops_g = eng.pack_g_operands()
    G = eng.general_w(operands=ops_g)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(steps):
        eng.general_w(operands=ops_g, out=G)
    ev[1].record()
    ev[6].record()
    for _ in range(steps):
        ops_p = eng.pack_g_operands()
    ev[7].record()
    del ops_p
    F = torch.empty((U, (M + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :M]
    lams = np.linspace(0.0, 1.0, steps + warmup)
    for lam in lams[:warmup]:
        eng.recommend(float(lam), 20, F_out=F)
    torch.cuda.synchronize()
    ev[2].record()
    for lam in lams[warmup:]:
        eng.recommend(float(lam), 20, F_out=F)
    ev[3].record()
    # F GEMM alone
    ev[4].record()
    for _ in range(steps):
        eng.resource(out=F)
    ev[5].record()
    torch.cuda.synchronize()
    # the same lambda step with the top-k selected inside the F-GEMM epilogue (F never written), and that kernel alone
    evf = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for lam in lams[:warmup]:
        eng.recommend(float(lam), 20, fused=True)
    torch.cuda.synchronize()
    evf[0].record()
    for lam in lams[warmup:]:
        eng.recommend(float(lam), 20, fused=True)
    evf[1].record()
    evf[2].record()
    for _ in range(steps):
        eng.resource_topk(20)
    evf[3].record()
    torch.cuda.synchronize()
    t_step_fused = evf[0].elapsed_time(evf[1]) / steps * 1e-3
    t_ftopk = evf[2].elapsed_time(evf[3]) / steps * 1e-3
    # F GEMM with W as a 24-bit per-column fixed point (3 digit planes): 1.33x fewer passes, error <= k_u 2^-25 s_j
    eng3 = ops.SpreadingEngine(d.n_users, d.n_items, users, items, w_mode="u8x3")
    eng3.G = eng.G
    eng3.scale(0.5)
    eng3.resource(out=F)
    torch.cuda.synchronize()
    evf[0].record()
    for _ in range(steps):
        eng3.resource(out=F)
    evf[1].record()
    torch.cuda.synchronize()
    t_f3 = evf[0].elapsed_time(evf[1]) / steps * 1e-3
    del eng3
    # lambda sweep as findLambda.py runs it: per lambda scale + F + filtered top-20 + the six metrics, one D2H at the end
    te = d.split()[2]
    test_pos = ops.seen_csr(torch.from_numpy(d.users[te]).to(dev), torch.from_numpy(d.items[te]).to(dev), U, M)
    eng.cooccurrence(operands=ops_g)
    sweep_l = np.linspace(0.0, 1.0, 11)
    eng.sweep(sweep_l[:2], 20, test_pos)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, sweep_res = eng.sweep(sweep_l, 20, test_pos)
    torch.cuda.synchronize()
    t_sweep = (time.perf_counter() - t0) / len(sweep_l)
    t_g = ev[0].elapsed_time(ev[1]) / steps * 1e-3
    t_pack = ev[6].elapsed_time(ev[7]) / steps * 1e-3
    t_step = ev[2].elapsed_time(ev[3]) / steps * 1e-3
    t_f = ev[4].elapsed_time(ev[5]) / steps * 1e-3
    _, peak_burst, _, how = peaks()
    flops = 2.0 * M * M * U
    try:
        e2e = spreading_e2e(d)
    except Exception as e:
        e2e = {"error": repr(e)[:300]}
    return {
        "e2e": e2e,
        "workload": f"hybrid spreading ml-1m shape ({U}x{M}, nnz(A)={sel.size}), top-20 full-rank filtered",
        "g_gemm": {"ms": round(t_g * 1e3, 4), "tflops": round(flops / t_g / 1e12, 2),
                   "kind": "G = A^T K_u^-1 A, u8 x4 digit planes of round(2^s/k_u), exact int32 accumulate; symmetric tile "
                           "schedule (tiles touching the upper triangle computed, mirrors stored) — tflops counts the full "
                           "matrix (SURVEY 8d)",
                   "operand_pack_ms": round(t_pack * 1e3, 4)},
        "f_gemm": {"ms": round(t_f * 1e3, 4), "tflops": round(flops / t_f / 1e12, 2),
                   "kind": "u8 x4 digit planes of per-column fixed-point W, exact int32 accumulate (w_mode u8x4)"},
        "f_gemm_u8x3": {"ms": round(t_f3 * 1e3, 4), "tflops": round(flops / t_f3 / 1e12, 2),
                        "frac_of_bf16_peak": round(flops / t_f3 / 1e12 / peak_burst, 4),
                        "kind": "opt-in w_mode u8x3: 24-bit per-column fixed point, 3 digit planes; |err| <= k_u 2^-25 s_j per entry "
                                "(passes the 1e-5 parity test at this shape, tests/test_gpu_fullsize.py), not the default because "
                                "the bound is not inside the tolerance for every possible input"},
        "lambda_step": {"ms": round(min(t_step, t_step_fused) * 1e3, 4), "users_per_s": round(U / min(t_step, t_step_fused), 1),
                        "what": "scale_w + F=A.W + filtered top-20, per lambda (faster of the two paths below; identical lists)",
                        "materialised_ms": round(t_step * 1e3, 4),
                        "materialised_what": "scale_w + F=A.W written (89 MB) + lgc_topk_rows",
                        "fused_ms": round(t_step_fused * 1e3, 4),
                        "fused_what": "scale_w + hs_resource_topk: top-20 selected in the F-GEMM epilogue, F never written",
                        "resource_topk_ms": round(t_ftopk * 1e3, 4),
                        "resource_topk_tflops": round(flops / t_ftopk / 1e12, 2)},
        "lambda_sweep": {"ms_per_lambda": round(t_sweep * 1e3, 4), "users_per_s": round(U / t_sweep, 1), "n_lambda": len(sweep_l),
                         "what": "findLambda.py pattern, wall clock: per lambda scale_w + F=A.W + filtered top-20 + P/R/F1/NDCG/H/I "
                                 "on the device (co-occurrence GEMM once), one device->host copy for the whole sweep",
                         "best": max(sweep_res, key=lambda m: m["precision"])},
        "roofline": {"bound": "tensor", "achieved": round(flops / t_f / 1e12, 2), "peak": peak_burst,
                     "unit": "TFLOP/s", "frac": round(flops / t_f / 1e12 / peak_burst, 4), "traffic": None,
                     "note": f"useful flops 2*U*M^2 of F=A.W (one pass counted; 4 int8 digit planes issued = 2 bf16-pass "
                             f"equivalents) / {how} bf16 dense peak"},
    }


def spreading_cpu_baseline(n_topk_users: int = 96):
    """The reference's own CPU path for BASELINE config 2 (ML-1M shape) on the box's host cores, through the oracle
    port of model/SpreadMethod/model.py (NumPy float64: np.dot x2, np.power, M^2 divide) and the literal per-user
    argsort + Python filter loop of recommend.py:35-47 on a bounded sample of users."""
    from oracle import spread_oracle as SO

    d = load_shape("ml-1m")
    tr, va, _ = d.split()
    sel = np.concatenate([tr, va])
    U, M = d.n_users, d.n_items
    A = SO.interaction_matrix(U, M, d.users[sel], d.items[sel])
    t0 = time.perf_counter(); G = SO.get_spreading_general_mat(A); t_g = time.perf_counter() - t0
    t0 = time.perf_counter(); W = SO.hybrids(A, G, 0.5); t_s = time.perf_counter() - t0
    t0 = time.perf_counter(); F = SO.get_resource(A, W); t_f = time.perf_counter() - t0
    seen = {}
    for u, i in zip(d.users[sel].tolist(), d.items[sel].tolist()):
        if u < n_topk_users:
            seen.setdefault(u, []).append(i)
    t0 = time.perf_counter(); SO.recommend_loop(F[:n_topk_users], seen, 20); t_k = (time.perf_counter() - t0) / n_topk_users
    flops = 2.0 * M * M * U
    step = t_s + t_f + t_k * U
    return {"kind": "port", "cores": torch.get_num_threads(), "dtype": "f64",
            "g_gemm_s": round(t_g, 3), "g_tflops": round(flops / t_g / 1e12, 3),
            "scale_s": round(t_s, 3), "f_gemm_s": round(t_f, 3), "f_tflops": round(flops / t_f / 1e12, 3),
            "topk_ms_per_user": round(t_k * 1e3, 2), "lambda_step_users_per_s": round(U / step, 1),
            "sample": f"G, HybridS and F once at full size; the argsort + Python filter loop on the first {n_topk_users} users, "
                      "extrapolated to all users for the lambda-step figure"}


def cpu_prop_csr_sample(adj: np.ndarray, n_users: int, n_items: int):
    """Best-effort CPU formulation of the same layer (not the reference's): normalised CSR built once, MKL CSR x dense."""
    ei = torch.from_numpy(adj)
    n = n_users + n_items
    deg = torch.bincount(ei[1], minlength=n).float()
    dinv = deg.pow(-0.5)
    dinv[torch.isinf(dinv)] = 0
    val = dinv[ei[0]] * dinv[ei[1]]
    A = torch.sparse_coo_tensor(torch.stack([ei[1], ei[0]]), val, (n, n)).coalesce().to_sparse_csr()
    torch.manual_seed(42)
    x = torch.empty(n, DIM).normal_(std=0.1)
    A @ x
    t0 = time.perf_counter()
    A @ x
    sec = time.perf_counter() - t0
    return prop_bytes(adj.shape[1], n, layers=1) / sec / 1e9, sec


def w_build_leg(d, dev, rank: int, world: int, steps: int = 3):
    """BASELINE config 5: hybrid W build on the ML-20M shape, sharded by item-column block with NO data-path
    collective: rank r computes G[:, J_r] = A^T K_u^-1 A[:, J_r] on its own tensor cores (exact int8 digit planes)."""
    from lgcnhs_b200 import ops

    tr, va, _ = d.split()
    sel = np.concatenate([tr, va])
    eng = ops.SpreadingEngine(d.n_users, d.n_items, torch.from_numpy(d.users[sel]).to(dev), torch.from_numpy(d.items[sel]).to(dev))
    U, M = d.n_users, d.n_items
    j0, j1 = M * rank // world, M * (rank + 1) // world
    operands = eng.pack_g_operands()
    # one GPU: the full matrix with the symmetric schedule (only the tiles touching the upper triangle are computed);
    # N GPUs: the SAME schedule dealt round-robin over the ranks, every tile and its mirror stored into every rank's G over
    # NVLink peer memory (fused GEMM + all-gather): each rank ends with the full matrix
    if world == 1:
        Gb = eng.general_w(operands=operands)
        run = lambda: eng.general_w(operands=operands, out=Gb)  # noqa: E731
    else:
        from lgcnhs_b200.dist import PeerGroup

        group = PeerGroup(dev)
        Gb, shared = eng.general_w_allgather(group, operands=operands)
        run = lambda: eng.general_w_allgather(group, operands=operands, shared=shared)  # noqa: E731
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    checksum = float(Gb.double().sum())     # mass conservation: sum_ij G[i,j] = sum_u k_u = nnz(A), on the FULL matrix
    symmetric = bool(torch.equal(Gb, Gb.T))
    if world > 1:                           # every rank holds the full matrix: all replicas must agree
        lo = torch.tensor([checksum], dtype=torch.float64, device=dev)
        hi = lo.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        symmetric = symmetric and bool(lo.item() == hi.item())
    flops = 2.0 * M * M * U
    _, peak_burst, peak_sus, how = peaks()
    tf = flops / (ms * 1e-3) / 1e12
    # share of the (256 x 64) output tiles the symmetric schedule actually computes
    tm, tn = (M + 255) // 256, (M + 63) // 64
    computed = sum(min(tm, ((n_ + 1) * 64 - 1) // 256 + 1) for n_ in range(tn)) / float(tm * tn)
    mcast = world > 1 and len(shared[1]) == 1
    del operands, Gb, eng, run
    shared = None
    torch.cuda.empty_cache()
    return {"workload": f"G = A^T K_u^-1 A on the ml-20m shape ({U}x{M}, nnz(A)={sel.size}), "
                        + ("symmetric tile schedule on 1 GPU" if world == 1 else
                           f"symmetric tile schedule dealt round-robin over {world} GPUs, tiles + mirrors stored into every "
                           "replica over NVLink (fused GEMM + all-gather; " + ("one NVSwitch multicast store per element"
                                                                              if mcast else "one store per peer")
                           + "), two device barriers"),
            "ms": round(ms, 3), "tflops": round(tf, 1), "scaling": "strong",
            "frac_of_bf16_peak": round(tf / (world * peak_sus), 4),
            "tiles_computed_frac": round(computed, 4), "tflops_issued": round(tf * computed, 1),
            "issued_frac_of_bf16_peak": round(tf * computed / (world * peak_sus), 4),
            "peak_note": f"tflops = useful 2*M^2*U flops of the FULL matrix (SURVEY 8d counts one pass and the full, not the "
                         f"symmetric-half, G) — it can exceed the dense peak because only tiles_computed_frac of the tiles are "
                         f"computed; tflops_issued counts the computed tiles only (4 int8 digit planes = 2 bf16-pass equivalents "
                         f"each); both / ({world} x {how} sustained bf16 peak)",
            "mass_check": {"sum_G": round(checksum, 3), "nnz_A": int(sel.size), "symmetric_and_replicas_equal": symmetric}}


def training_leg(dev, steps: int, warmup: int, rank: int = 0, world: int = 1):
    """BASELINE config 4: LightGCN 3-layer dim-64 BPR training step + full-rank eval, Amazon-Book shape, on `world` GPUs:
    the rows of A_hat are partitioned over the ranks for the forward and the gradient propagation (fused peer-store
    exchange), Adam runs on the rows a rank owns and pushes them to every replica, the evaluation is sharded by user
    block (FusedBPRTrainer(distributed=True), sharded_topk_layer0)."""
    from lgcnhs_b200 import ops
    from lgcnhs_b200.trainer import FusedBPRTrainer, sharded_topk_layer0
    from model.LightGCN.evaluation import _topk_layer0
    from model.LightGCN.model import LightGCN

    dist_on = world > 1
    barrier = (lambda: torch.distributed.barrier()) if dist_on else None
    d = load_shape("amazon-book", rank, barrier)
    adj_np, (tr, va, te) = train_adj(d)
    adj = torch.from_numpy(adj_np).to(dev)
    torch.manual_seed(42)
    model = LightGCN(d.n_users, d.n_items, DIM, K_LAYERS).to(dev)
    trainer = FusedBPRTrainer(model, adj, lr=1e-3, eps_reg=1e-6, distributed=dist_on)
    B = 1024
    from model.LightGCN.loss import sampleMiniBatch

    # one reference iteration (train.py:125-144): sample a mini-batch (device negative sampler), forward, BPR,
    # backward, Adam.  Every rank draws the same mini-batch (same seed).
    train_ei = torch.from_numpy(np.stack([d.users[tr], d.items[tr]])).to(dev)
    torch.manual_seed(42)

    def one_step():
        u, p, n = sampleMiniBatch(B, train_ei)
        return trainer.step(u, p, n)

    def sync():
        torch.cuda.synchronize()
        if dist_on:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not dist_on:
            return x
        t = torch.tensor([x], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(warmup, 3)):
        one_step()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = one_step()
    e1.record()
    sync()
    ms = max_over_ranks(e0.elapsed_time(e1) / steps)
    parity = None
    if dist_on:
        # all replicas of the weight table must be bit-identical (every row is written by exactly one owner)
        chk = trainer.X0.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        parity = {"weight_replicas_identical": bool(lo.item() == hi.item())}
    e_tr = torch.from_numpy(np.stack([d.users[tr], d.items[tr]]))
    if dist_on:
        seen = ops.seen_csr(e_tr[0].to(dev), e_tr[1].to(dev), d.n_users, d.n_items)
        full, mine = sharded_topk_layer0(model, d.n_users, d.n_items, seen, 20, rank, world)
        sync()
        e0.record()
        full, mine = sharded_topk_layer0(model, d.n_users, d.n_items, seen, 20, rank, world)
        e1.record()
        sync()
        ms_eval = max_over_ranks(e0.elapsed_time(e1))
        ref = _topk_layer0(model, d.n_users, d.n_items, [e_tr], 20)      # the same ranking on one GPU
        parity["eval_lists_equal_1gpu"] = bool(torch.equal(ref, full))
    else:
        _topk_layer0(model, d.n_users, d.n_items, [e_tr], 20)
        torch.cuda.synchronize()
        e0.record()
        _topk_layer0(model, d.n_users, d.n_items, [e_tr], 20)
        e1.record()
        torch.cuda.synchronize()
        ms_eval = e0.elapsed_time(e1)
    if dist_on:
        return {"workload": f"LightGCN K=3 D=64 BPR step (batch {B}) + full-rank top-20 eval, amazon-book shape "
                            f"(U={d.n_users}, M={d.n_items}, nnz={adj_np.shape[1]}) on {world} GPUs",
                "step_ms": round(ms, 4), "loss": round(float(loss[0]), 5), "scaling": "strong",
                "eval_ms": round(ms_eval, 3), "eval_users_per_s": round(d.n_users / (ms_eval * 1e-3), 1),
                "parity": parity,
                "what": "rows of A_hat partitioned by nnz (users and items separately) for the forward and the gradient "
                        "propagation with the fused peer-store exchange + device barrier per layer; BPR replicated (same "
                        "batch); Adam on owned rows, new rows pushed to every replica over NVLink; eval sharded by user block, "
                        "ids all-gathered; device time, max over ranks"}
    # e2e of the full-rank recommendation through the reference-facing call: recommendForAllUser(model, U, M, train_adj,
    # val_adj, test_adj, k) with the HOST adjacency tensors buildGraph returns -> dict{uid: [k ids]} (+ np.save)
    try:
        install_cfg("LightGCN", k=20)
        from model.LightGCN.recommend import recommendForAllUser

        adj_host = torch.from_numpy(adj_np)
        val_host = torch.from_numpy(bipartite(d, va))
        test_host = torch.from_numpy(bipartite(d, te))
        out = {}

        def call():
            out["rec"] = recommendForAllUser(model, d.n_users, d.n_items, adj_host, val_host, test_host, 20)
        ms_e2e = wall_ms(call, reps=3)
        eval_e2e = {"ms": round(ms_e2e, 2), "users_per_s": round(d.n_users / (ms_e2e * 1e-3), 1), "unit": "users/s",
                    "h2d_bytes_per_step": int((adj_host.numel() + val_host.numel() + test_host.numel()) * 8),
                    "d2h_bytes_per_step": int(d.n_users * 20 * 8),
                    "what": "recommendForAllUser(model, U, M, train_adj, val_adj, test_adj, 20): host adjacency tensors -> "
                            "dict{uid: top-20}; wall clock incl. upload, adjacency -> edge list, mask CSR, fused score/top-k "
                            "kernel, download, dict + np.save (median of 3)"}
    except Exception as e:
        eval_e2e = {"error": repr(e)[:300]}
    hbm = peaks()[0]
    gbs = trainer.step_bytes(B) / (ms * 1e-3) / 1e9
    gbs_c = trainer.step_bytes_compulsory(B) / (ms * 1e-3) / 1e9
    return {"workload": f"LightGCN K=3 D=64 BPR step (batch {B}) + full-rank top-20 eval, amazon-book shape "
                        f"(U={d.n_users}, M={d.n_items}, nnz={adj_np.shape[1]})",
            "step_ms": round(ms, 4), "loss": round(float(loss[0]), 5),
            "step_compulsory_gbs": round(gbs_c, 1), "step_frac_of_hbm_peak": round(gbs_c / hbm, 4),
            "step_no_reuse_gbs": round(gbs, 1),
            "what": "device mini-batch + negative sampling kernel, then ONE CUDA graph per step: 2K fused SpMM layers (fwd + grad) + fused BPR fwd/bwd scatter + Adam (device-resident bias corrections); no host sync",
            "eval_ms": round(ms_eval, 2), "eval_users_per_s": round(d.n_users / (ms_eval * 1e-3), 1), "eval_e2e": eval_e2e,
            "eval_what": "ONE fused kernel + a candidate merge: layer-0 score tiles on the tensor cores (tcgen05 kind::tf32, 3xTF32 split) + train-pair fill(-1024) + top-20 over all 91 599 items; the U x M score matrix is never written; the mask CSR of the train pairs is built once per graph (2nd evaluation timed)"}


_REAL_STDOUT = None


def _capture_stdout():
    """The contract is ONE JSON line on stdout; libraries (NCCL's version banner, torchrun notices) also
    write to fd 1.  Point fd 1 at stderr for the whole run and keep the real stdout for the JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="ml-20m")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spreading", action="store_true")
    ap.add_argument("--split", type=int, default=1, help="multi-GPU: partition user rows and item rows separately (1, default: every rank gets the same mix of both row classes, one mixed launch per layer) or as one range (0)")
    ap.add_argument("--mode", default="p2p", choices=["p2p", "p2p-nccl", "nccl"], help="multi-GPU layer exchange")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    from lgcnhs_b200 import _lib, ops
    from lgcnhs_b200.dist import RowPartitionedPropagation, init_dist

    rank, world, local = init_dist()
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    barrier = (lambda: torch.distributed.barrier()) if world > 1 else None
    clk = Clocks(local)   # started early: nvidia-smi needs seconds to enumerate an 8-GPU box
    d = load_shape(args.shape, rank, barrier)
    adj_np, _ = train_adj(d)
    n = d.n_users + d.n_items
    nnz = int(adj_np.shape[1])
    adj = torch.from_numpy(adj_np).to(dev)
    torch.manual_seed(42)
    users_w = torch.empty(d.n_users, DIM).normal_(std=0.1)
    items_w = torch.empty(d.n_items, DIM).normal_(std=0.1)
    x0_host = torch.cat([users_w, items_w]).pin_memory()
    x0 = x0_host.to(dev)

    if world == 1:
        g = ops.NormGraph(adj, n)
        E = torch.empty_like(x0)
        tmp = (torch.empty_like(x0), torch.empty_like(x0))
        step = lambda: g.propagate_mean(x0, K_LAYERS, out=E, tmp=tmp)  # noqa: E731
        launches_per_step = K_LAYERS
    else:
        prop = RowPartitionedPropagation(adj, n, DIM, mode=args.mode, split=d.n_users if args.split else None)
        step = lambda: prop.propagate_mean(x0, K_LAYERS)  # noqa: E731
        launches_per_step = K_LAYERS

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    _lib.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    clk.begin()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    clk.end()
    launches = _lib.launch_count()      # our kernels launched inside the timed region
    if clk.t1 - clk.t0 < 0.5:
        clk.probe(step, sync_all)
    clk.stop()
    ms = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    bytes_step = prop_bytes(nnz, n)
    value = bytes_step / (ms * 1e-3) / 1e9
    parity_vs_1gpu = None
    if world > 1:
        # every rank recomputes the K-layer mean on ONE GPU and compares it with the replica the partitioned path left
        # on it: the multi-GPU result must equal the single-GPU one to fp32 round-off (a row's summation path depends on
        # the launch it is part of), on every rank
        ref = ops.NormGraph(adj, n).propagate_mean(x0, K_LAYERS)
        got = prop.propagate_mean(x0, K_LAYERS)
        rel = ((got - ref).abs().max() / ref.abs().max()).reshape(1)
        torch.distributed.all_reduce(rel, op=torch.distributed.ReduceOp.MAX)
        parity_vs_1gpu = {"max_abs_diff_over_max_abs": float(rel.item()), "tolerance": 2e-6,
                          "ok": bool(rel.item() <= 2e-6), "what": "K-layer mean on every rank vs the same call on one GPU"}
        del ref, got

    line = None
    if rank == 0:
        hbm, _, _, how = peaks()
        line = {
            "metric": "LightGCN prop GB/s", "value": round(value, 2), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 5), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"LightGCN K={K_LAYERS} D={DIM} fused propagation + layer mean, {args.shape} shape "
                                   f"train graph (U={d.n_users}, M={d.n_items}, nnz={nnz}, N={n})",
                       "l2": "inputs larger than L2 (CSR stream 8*nnz = %.0f MB per layer; X 42 MB is L2-resident by design)"
                             % (8 * nnz / 1e6),
                       "parallelism": "1 GPU" if world == 1 else
                       f"row partition by nnz over {world} GPUs (users and items separately: {bool(args.split)}), one mixed launch per "
                       f"layer and rank, {args.mode} exchange fused into the SpMM: "
                       + ("every row stored ONCE to an NVSwitch multicast address (NVLS replicates it into all replicas)"
                          if getattr(prop, "mcast", None) is not None else
                          "every row stored into each peer's replica (CUDA IPC); multicast unavailable: "
                          + str(getattr(prop, "mcast_error", "disabled"))[:160])
                       + ", one device barrier per layer"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        per_layer_ms = ms / K_LAYERS
        compulsory = nnz * 8 + (n + 1) * 4 + 2 * n * 4 * DIM
        comp_gbs = compulsory / (per_layer_ms * 1e-3) / 1e9
        gather_gbs = nnz * 4 * DIM / (per_layer_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(f"spmm_layer_kernel/{args.shape}")
        line["roofline"] = {
            "bound": "hbm", "achieved": round(comp_gbs, 2), "peak": hbm, "unit": "GB/s",
            "frac": round(comp_gbs / hbm / world, 4),
            "traffic": traffic, "traffic_source": traffic_src, "kernel": "spmm_layer_kernel<64,NPEER,UN>",
            "us_per_layer": round(per_layer_ms * 1e3, 2),
            "bytes_model": "COMPULSORY bytes per layer = nnz*8 (colidx + val, streamed once) + (N+1)*4 (rowptr) + 2*N*4*D "
                           "(X read once, Y written once): every byte the layer must move through HBM, a true <= 1 bound "
                           f"against the {how} HBM copy bandwidth (x n_gpus)",
            "no_reuse_gbs": round(value, 2), "no_reuse_frac": round(value / hbm / world, 4),
            "no_reuse_note": "`value` uses SURVEY.md 8d's no-reuse model (264 B/nnz + 260 B/node per layer: one 256-B row "
                             "gather per non-zero).  X (42 MB) is L2-resident, so those gathers are served by L2 and "
                             "no_reuse_frac is NOT an HBM fraction (it exceeds 1); the gathers have their own bound below",
        }
        # L2 gather bound: the 256-B row gathers (nnz * 4D bytes per layer) against this GPU's measured throughput for
        # uniformly random whole-row reads from a table of X's size (lgc_probe_gather, same LDG.128 x 16-lane pattern)
        try:
            pr = ops.probe_gather_gbs(n, DIM, n_gathers=nnz, device=dev)
            line["roofline"]["l2_gather"] = {
                "achieved": round(gather_gbs, 1), "peak": round(pr["gbs"], 1), "unit": "GB/s",
                "frac": round(gather_gbs / pr["gbs"] / world, 4),
                "bytes_model": "nnz * 4 * D gathered row bytes per layer / time",
                "probe": {"what": "lgc_probe_gather: uniformly random whole-row reads (LDG.128 x 16 lanes) from an (N, 64) fp32 "
                                  "table, 64 warps/SM, 8 gathers in flight per lane, measured in this run (best of 5)",
                          "table_mb": round(pr["table_mb"], 1), "rows": pr["rows"], "us": round(pr["us"], 1)}}
        except Exception as e:
            line["roofline"]["l2_gather"] = {"error": repr(e)[:200]}
        if world > 1:
            line["parity_vs_1gpu"] = parity_vs_1gpu

    # ---- e2e through the reference-facing module call, host buffers, N GPUs ----
    from model.LightGCN.model import LightGCN

    if world == 1:
        # (a) the module call, one stream: upload, forward, download strictly in sequence
        model = LightGCN(d.n_users, d.n_items, DIM, K_LAYERS).to(dev)
        out_host = torch.empty((n, DIM), dtype=torch.float32).pin_memory()
        with torch.no_grad():
            def e2e_step():
                model.users_emb.weight.copy_(x0_host[: d.n_users], non_blocking=True)
                model.items_emb.weight.copy_(x0_host[d.n_users:], non_blocking=True)
                uf, _, itf, _ = model.forward(adj)
                out_host[: d.n_users].copy_(uf, non_blocking=True)
                out_host[d.n_users:].copy_(itf, non_blocking=True)
            for _ in range(3):
                e2e_step()
            sync_all()
            ev0.record()
            n_e2e = max(3, args.steps // 2)
            for _ in range(n_e2e):
                e2e_step()
            ev1.record()
            sync_all()
        ms_e2e_serial = ev0.elapsed_time(ev1) / n_e2e
        # (b) the same work through the pipelined host-table entry point: upload / K layers / download on three streams,
        # consecutive calls overlap; every call still moves its own 42 MB in and 42 MB out inside the timed region
        from lgcnhs_b200.propagation import PipelinedPropagation

        pipe = PipelinedPropagation(adj, n, DIM, K_LAYERS)
        outs = [torch.empty((n, DIM), dtype=torch.float32).pin_memory() for _ in range(2)]
        for i in range(4):
            pipe.submit(x0_host, outs[i & 1])
        pipe.synchronize()
        torch.cuda.synchronize()
        n_e2e = max(6, args.steps)
        t0 = time.perf_counter()
        for i in range(n_e2e):
            pipe.submit(x0_host, outs[i & 1])
        pipe.synchronize()
        ms_e2e = (time.perf_counter() - t0) * 1e3 / n_e2e
        e2e_check = float((outs[(n_e2e - 1) & 1] - out_host).abs().max())      # same E as the module call
    else:
        # N GPUs: every rank uploads ONE slice of e^0 over its own PCIe link, the slices are all-gathered over NVLink,
        # and every rank downloads only the rows it computed — the job moves N*D*4 bytes each way, like one GPU
        chunk = (n + world - 1) // world
        x0_pad = torch.zeros((world * chunk, DIM), dtype=torch.float32, device=dev)
        c0, c1 = rank * chunk, min((rank + 1) * chunk, n)
        out_host = torch.empty((n, DIM), dtype=torch.float32).pin_memory()
        my_rows = [(a, b) for a, b in prop.parts[rank] if b > a]

        def e2e_step():
            if c1 > c0:
                x0_pad[c0:c1].copy_(x0_host[c0:c1], non_blocking=True)
            torch.distributed.all_gather_into_tensor(x0_pad, x0_pad[rank * chunk:(rank + 1) * chunk])
            E = prop.propagate_mean(x0_pad[:n], K_LAYERS)
            for a, b in my_rows:
                out_host[a:b].copy_(E[a:b], non_blocking=True)
        for _ in range(3):
            e2e_step()
        sync_all()
        ev0.record()
        n_e2e = max(3, args.steps // 2)
        for _ in range(n_e2e):
            e2e_step()
        ev1.record()
        sync_all()
        t = torch.tensor([ev0.elapsed_time(ev1) / n_e2e], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_e2e = float(t.item())
    if rank == 0:
        line["e2e"] = {"value": round(bytes_step / (ms_e2e * 1e-3) / 1e9, 2), "unit": "GB/s",
                       "ms_per_step": round(ms_e2e, 4), "h2d_bytes_per_step": int(x0_host.numel() * 4),
                       "d2h_bytes_per_step": int(out_host.numel() * 4),
                       "what": ("PipelinedPropagation.submit(pinned-host e^0, pinned-host E): upload / K fused layers / download on "
                                "three streams, consecutive calls overlap (wall clock over the whole loop incl. the final drain); "
                                "every call moves its own tables"
                                if world == 1 else
                                "pinned-host e^0 uploaded in per-rank slices + NVLink all-gather, K fused layers with fused row "
                                "exchange, every rank downloads the rows it computed")}

    if rank == 0 and world == 1:
        line["e2e"]["serial_ms_per_step"] = round(ms_e2e_serial, 4)
        line["e2e"]["serial_what"] = "LightGCN.forward(edge_index) on ONE stream: upload, K layers, download in sequence"
        line["e2e"]["max_abs_diff_vs_module_call"] = e2e_check
    # ---- the other two figures of the metric: W TFLOP/s and top-20 users/s (config 2), rank 0 ----
    if rank == 0 and not args.no_spreading:
        try:
            line["spreading"] = spreading_leg(dev, steps=max(3, min(args.steps, 10)), warmup=3)
        except Exception as e:  # keep the primary line even if the secondary leg fails
            line["spreading"] = {"error": repr(e)[:300]}
    if not args.no_spreading and args.shape == "ml-20m":
        try:
            wb = w_build_leg(d, dev, rank, world)
        except Exception as e:
            wb = {"error": repr(e)[:300]}
        if rank == 0:
            line["w_build"] = wb
    if world > 1 and not args.no_spreading:
        try:
            tl = training_leg(dev, steps=max(5, min(args.steps, 20)), warmup=3, rank=rank, world=world)
        except Exception as e:
            tl = {"error": repr(e)[:300]}
        if rank == 0:
            line["training"] = tl
    if rank == 0 and world == 1 and not args.no_spreading:
        try:
            line["training"] = training_leg(dev, steps=max(5, min(args.steps, 20)), warmup=3)
        except Exception as e:
            line["training"] = {"error": repr(e)[:300]}
        try:
            line["fusion"] = fusion_leg(dev)
        except Exception as e:
            line["fusion"] = {"error": repr(e)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gbs, sec = cpu_prop_sample(adj_np, d.n_users, d.n_items, steps=1, warmup=1)
        line["cpu_baseline"] = {"value": round(gbs, 3), "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "1 of 3 propagation layers over the full graph (PyG-equivalent oracle port: "
                                          "gcn_norm + index_select + scatter_add); %.2f s per layer" % sec}
        try:
            gbs_c, sec_c = cpu_prop_csr_sample(adj_np, d.n_users, d.n_items)
            line["cpu_baseline"]["best_effort_csr"] = {
                "value": round(gbs_c, 3), "unit": "GB/s",
                "what": "same layer as MKL CSR x dense with the normalised CSR prebuilt (not the reference's formulation); "
                        "%.3f s per layer" % sec_c}
        except Exception as e:
            line["cpu_baseline"]["best_effort_csr"] = {"error": repr(e)[:200]}
        if "spreading" in line and "error" not in line["spreading"]:
            try:
                line["spreading"]["cpu_baseline"] = spreading_cpu_baseline()
            except Exception as e:
                line["spreading"]["cpu_baseline"] = {"error": repr(e)[:200]}
    if rank == 0:
        emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
