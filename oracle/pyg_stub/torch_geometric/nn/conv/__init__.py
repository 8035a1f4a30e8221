"""MessagePassing stub: aggr='add', flow='source_to_target', dense edge_index path of PyG 2.6.1."""
import torch
from torch import nn


class MessagePassing(nn.Module):
    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, x, norm):
        row, col = edge_index[0], edge_index[1]
        msg = self.message(x_j=x.index_select(0, row), norm=norm)       # __collect__ + message
        out = torch.zeros_like(x)
        return out.scatter_add_(0, col.view(-1, 1).expand(-1, x.shape[1]), msg)   # aggregate (sum over targets)

    def message(self, x_j, **kwargs):
        return x_j
