"""Drop-in for /root/reference/model/LightGCN/model.py — same class name, constructor,
attributes (`users_emb`, `items_emb`) and `forward(edge_index)` 4-tuple, but the propagation
runs on hand-written sm_100a kernels (no torch_geometric)."""
import torch
from torch import nn

from lgcnhs_b200.propagation import lightgcn_forward


class LightGCN(nn.Module):
    def __init__(self, user_num: int, item_num: int, embedding_dim: int, layers: int) -> None:
        super().__init__()
        self.user_num = user_num
        self.item_num = item_num
        self.embedding_dim = embedding_dim
        self.layers = layers
        # e_u^0 (user_num, dim) and e_i^0 (item_num, dim), N(0, 0.1^2) — reference model.py:31-38
        self.users_emb = nn.Embedding(num_embeddings=self.user_num, embedding_dim=self.embedding_dim)
        self.items_emb = nn.Embedding(num_embeddings=self.item_num, embedding_dim=self.embedding_dim)
        nn.init.normal_(self.users_emb.weight, std=0.1)
        nn.init.normal_(self.items_emb.weight, std=0.1)

    def forward(self, edge_index: torch.Tensor) -> tuple:
        """edge_index: the (2, 2E) symmetric adjacency COO of utils/graph.py.
        Returns (e_u^K-mean, e_u^0, e_i^K-mean, e_i^0) exactly like reference model.py:74."""
        return lightgcn_forward(self.users_emb.weight, self.items_emb.weight, edge_index, self.layers)
