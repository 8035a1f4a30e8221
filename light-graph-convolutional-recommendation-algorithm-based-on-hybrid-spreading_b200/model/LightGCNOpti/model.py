"""Drop-in for /root/reference/model/LightGCNOpti/model.py: LightGCN whose e^0 is initialised
from side features through a Linear layer (reference model.py:36-49); same forward."""
import torch
from torch import nn

from lgcnhs_b200.propagation import lightgcn_forward


class LightGCNOpti(nn.Module):
    def __init__(self, user_num: int, item_num: int, embedding_dim: int, layers: int,
                 user_features: torch.Tensor, item_features: torch.Tensor) -> None:
        super().__init__()
        self.user_num = user_num
        self.item_num = item_num
        self.embedding_dim = embedding_dim
        self.layers = layers
        self.user_linear = nn.Linear(user_features.size(1), embedding_dim)
        self.item_linear = nn.Linear(item_features.size(1), embedding_dim)
        user_emb_init = self.user_linear(user_features)
        item_emb_init = self.item_linear(item_features)
        self.users_emb = nn.Embedding(num_embeddings=self.user_num, embedding_dim=self.embedding_dim)
        self.users_emb.weight = nn.Parameter(user_emb_init)
        self.items_emb = nn.Embedding(num_embeddings=self.item_num, embedding_dim=self.embedding_dim)
        self.items_emb.weight = nn.Parameter(item_emb_init)

    def forward(self, edge_index: torch.Tensor) -> tuple:
        return lightgcn_forward(self.users_emb.weight, self.items_emb.weight, edge_index, self.layers)
