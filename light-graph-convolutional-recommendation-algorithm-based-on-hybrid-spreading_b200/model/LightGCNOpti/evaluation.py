"""Drop-in for /root/reference/model/LightGCNOpti/evaluation.py (differs from LightGCN/evaluation.py
only in names): same getValRecommendations / calValLoss, one implementation."""
from model.LightGCN.evaluation import calValLoss, getValRecommendations  # noqa: F401
