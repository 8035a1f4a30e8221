"""Drop-in for /root/reference/model/SpreadMethod/model.py — NumPy float64 in, NumPy float64 out,
same five functions; the two dense contractions run on the B200 tensor cores (tcgen05 / TMEM) and
the degree scaling as one fused pass.

These array-level entry points exist for callers that hold NumPy matrices (findLambda.py,
SpreadLightGCN*/model.py); every call pays host<->device copies of its operands.  The
recommenders in model/SpreadMethod/recommend.py use lgcnhs_b200.ops.SpreadingEngine directly and
stay on the device from the interaction list to the top-k ids."""
import numpy as np
import torch

from lgcnhs_b200 import ops
from utils.log import logger
from utils.wrapper import calTimes


def _dev() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("model.SpreadMethod: no CUDA device - the B200 drop-in has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _engine_from_dense(A: np.ndarray) -> ops.SpreadingEngine:
    """Interaction list of a dense 0/1 matrix -> device engine (cached on the array's identity)."""
    key = (A.__array_interface__["data"][0], A.shape)
    hit = _engine_from_dense.cache.get(key)
    if hit is not None and hit[1] is A:
        return hit[0]
    if not np.all((A == 0) | (A == 1)):
        raise ValueError("interaction matrix must be binary (reference utils/trans.py:13-29 writes only 0/1)")
    u, i = np.nonzero(A)
    dev = _dev()
    eng = ops.SpreadingEngine(A.shape[0], A.shape[1], torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev))
    _engine_from_dense.cache.clear()
    _engine_from_dense.cache[key] = (eng, A)
    return eng


_engine_from_dense.cache = {}


@calTimes(logger, "通用扩散矩阵计算完成")
def getSpreadingGeneralMat(A: np.ndarray) -> np.ndarray:
    """general_W = (A^T / k_u) . A   (reference model.py:14-27)."""
    eng = _engine_from_dense(A)
    return eng.general_w().double().cpu().numpy()


def _scale(A: np.ndarray, general_W: np.ndarray, lam: float):
    eng = _engine_from_dense(A)
    G = torch.from_numpy(np.ascontiguousarray(general_W, dtype=np.float32)).to(eng.dev)
    return eng, eng.scale(lam, G=G, want_w32=True)


@calTimes(logger, "扩散资源矩阵计算完成")
def ProbS(A: np.ndarray, general_W: np.ndarray) -> np.ndarray:
    """W = general_W / k_j   (reference model.py:30-43) == HybridS with lambda = 1."""
    return _scale(A, general_W, 1.0)[1].double().cpu().numpy()


@calTimes(logger, "扩散资源矩阵计算完成")
def HeatS(A: np.ndarray, general_W: np.ndarray) -> np.ndarray:
    """W = general_W / k_i   (reference model.py:46-60) == HybridS with lambda = 0."""
    return _scale(A, general_W, 0.0)[1].double().cpu().numpy()


@calTimes(logger, "扩散资源矩阵计算完成")
def HybridS(A: np.ndarray, general_W: np.ndarray, Lambda: float) -> np.ndarray:
    """W = general_W / (k_i^(1-Lambda) k_j^Lambda), zero denominators -> 1   (reference model.py:63-85)."""
    return _scale(A, general_W, float(Lambda))[1].double().cpu().numpy()


@calTimes(logger, "资源矩阵计算完成")
def getResource(A: np.ndarray, W: np.ndarray) -> np.ndarray:
    """F_new = A . W   (reference model.py:88-99): W is split into bf16 hi/mid/lo planes on the device."""
    eng = _engine_from_dense(A)
    M = eng.M
    Wd = torch.from_numpy(np.ascontiguousarray(W, dtype=np.float32)).to(eng.dev)
    if float(Wd.min()) < 0.0:
        raise ValueError("getResource: W must be non-negative (transfer weights are), got negative entries")
    # identity scaling (unit degrees) re-uses the fused scale kernel to transpose W and emit the operand planes
    ki = eng.ki
    eng.ki = torch.ones(M, dtype=torch.int32, device=eng.dev)
    try:
        eng.scale(0.0, G=Wd)
    finally:
        eng.ki = ki
    return eng.resource().double().cpu().numpy()
