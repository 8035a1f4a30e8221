"""Drop-in for /root/reference/model/LightGCNOpti/loss.py (byte-identical to LightGCN/loss.py in
the reference): same BPRLoss / sampleMiniBatch, one implementation."""
from model.LightGCN.loss import BPRLoss, sampleMiniBatch  # noqa: F401
