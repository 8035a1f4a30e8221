"""CPU: the C-ABI library builds/loads without a GPU and exports every symbol include/lgcnhs.h declares."""
import pytest


def test_build_and_symbols():
    import __graft_entry__ as g

    g.build()
    from lgcnhs_b200 import _lib

    L = _lib.lib()
    syms = _lib.header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/lgcnhs.h but not exported"
        assert s in _lib._SIGS, f"{s} has no ctypes signature"
    assert set(_lib._SIGS) == set(syms)
    assert L.lgc_abi_version() == 1


def test_argument_errors_are_reported_not_thrown(lib):
    """Validation happens before any CUDA call: usable without a device, errors come back as codes + message."""
    from lgcnhs_b200._lib import LgcnhsError, check

    rc = lib.hs_gemm_planes(7, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1.0, 0)
    assert rc == -1 and b"kind" in lib.lgc_last_error_string()
    rc = lib.lgc_topk_rows(0, 1, 1, 1, 0, 0, 0, 1, 0, 0, 0)  # null score pointer
    assert rc == -1
    with pytest.raises(LgcnhsError):
        check(lib.lgc_adam_step(0, 0, 0, 0, 0, 0.1, 0.9, 0.999, 1e-8, 0.1, 0.1, 0), "adam")
    assert lib.lgc_csr_max_chunks(1 << 20) == (1 << 20) // 1024 + (1 << 20) // 256 + 1
    assert lib.lgc_bpr_scratch_floats(1024) >= 2 + 2 * 1024


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing on the host."""
    import torch

    from lgcnhs_b200._lib import LgcnhsError
    from lgcnhs_b200.ops import NormGraph
    from lgcnhs_b200.propagation import lightgcn_forward

    with pytest.raises(LgcnhsError):
        NormGraph(torch.zeros((2, 4), dtype=torch.int64), 4)
    with pytest.raises(RuntimeError):
        lightgcn_forward(torch.zeros(2, 64), torch.zeros(2, 64), torch.zeros((2, 4), dtype=torch.int64), 3)
    if not torch.cuda.is_available():
        # the metrics drop-in (consumer of the top-k lists) has no host implementation either
        import _stub_const
        import numpy as np

        _stub_const.install()
        from metrics.accurate import getAccurateMetrics
        from metrics.diversity import getDiversityMetrics

        rec = torch.zeros((2, 3), dtype=torch.long)
        with pytest.raises(RuntimeError):
            getAccurateMetrics({0: [1]}, rec, 3)
        with pytest.raises(RuntimeError):
            getDiversityMetrics(rec, {0: 1}, np.zeros((2, 4)), 3)


def test_product_path_never_imports_the_oracle():
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "light-graph-convolutional-recommendation-algorithm-based-on-hybrid-spreading_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} reaches into oracle/"


def test_ctypes_signatures_match_header_arity():
    """Every binding in _lib._SIGS takes exactly as many arguments as its declaration in include/lgcnhs.h (a
    mismatch would corrupt the call frame silently: ctypes cannot check it)."""
    import re

    from lgcnhs_b200 import _lib

    txt = open(_lib.HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for name, (_, args) in _lib._SIGS.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", txt, flags=re.S)
        assert m, f"{name}: declaration not found"
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert n == len(args), f"{name}: header declares {n} parameters, ctypes binding has {len(args)}"
    # new argument checks that need no device
    lib = _lib.lib()
    assert lib.lgc_score_topk(0, 0, 0, 1, 1, 64, 0, 0, 0.0, 0, 0, 0, 1, 0, 0, 0) == -1       # null embeddings
    assert lib.lgc_metrics_topk(0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0) == -1                      # null lists
    assert lib.lgc_peer_barrier(0, None, 0, 1, 1, 0) == -1
    assert lib.lgc_spmm_long_row(100) == -1 and lib.lgc_spmm_long_row(0) == 0
    assert lib.lgc_score_topk_config(300) == -1 and lib.lgc_score_topk_config(512) == 0
    assert lib.lgc_metrics_scratch_bytes(1000) >= 4000
