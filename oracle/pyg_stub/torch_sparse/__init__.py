"""SparseTensor stub: only what utils/graph.py:46-47 uses (row/col/sparse_sizes -> to_dense, fp32 ones)."""
import torch


class SparseTensor:
    def __init__(self, row, col, sparse_sizes, value=None):
        self.row, self.col, self.sizes = row, col, sparse_sizes

    def to_dense(self):
        out = torch.zeros(self.sizes, dtype=torch.float32)
        out[self.row, self.col] = 1.0
        return out
