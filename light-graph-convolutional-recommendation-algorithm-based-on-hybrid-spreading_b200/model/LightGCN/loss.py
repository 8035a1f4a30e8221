"""Drop-in for /root/reference/model/LightGCN/loss.py: BPRLoss and sampleMiniBatch with the
reference's signatures.  BPRLoss runs as ONE fused CUDA kernel forward and one backward
(lgc_bpr_rows) instead of ~12 elementwise/reduce launches."""
import torch

from lgcnhs_b200 import ops
from lgcnhs_b200.sampling import structured_negative_sampling


class _BPRRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u_f, u_0, p_f, p_0, n_f, n_0, lambda_val):
        rows = [t.contiguous() for t in (u_f, u_0, p_f, p_0, n_f, n_0)]
        ctx.save_for_backward(*rows)
        ctx.lambda_val = float(lambda_val)
        return ops.bpr_rows(rows, ctx.lambda_val)[0]

    @staticmethod
    def backward(ctx, grad_out):
        rows = ctx.saved_tensors
        grads = [torch.empty_like(t) for t in rows]
        ops.bpr_rows(rows, ctx.lambda_val, grads)
        return tuple(g * grad_out for g in grads) + (None,)


def BPRLoss(users_emb_final: torch.Tensor, users_emb_0: torch.Tensor,
            pos_items_emb_final: torch.Tensor, pos_items_emb_0: torch.Tensor,
            neg_items_emb_final: torch.Tensor, neg_items_emb_0: torch.Tensor,
            lambda_val: float) -> torch.Tensor:
    """loss = -mean(softplus(s+ - s-)) + lambda * (|u0|^2 + |p0|^2 + |n0|^2)   (reference loss.py:29-42,
    sign and un-normalised L2 term kept as they are)."""
    return _BPRRows.apply(users_emb_final, users_emb_0, pos_items_emb_final, pos_items_emb_0,
                          neg_items_emb_final, neg_items_emb_0, lambda_val)


def sampleMiniBatch(batch_size: int, edge_index: torch.Tensor) -> tuple:
    """batch_size (user, pos, neg) triplets (reference loss.py:46-70).  The reference negative-samples
    ALL edges and then keeps batch_size of them with `random.choices` (uniform, with replacement, Python RNG,
    unseeded); picking the rows first and sampling negatives only for those is the same distribution at 1/E-th
    of the work.  The row indices are drawn on the device (torch.randint: uniform with replacement, like
    random.choices) so that the step never synchronises with the host."""
    n_edges = edge_index.shape[1]
    indices = torch.randint(n_edges, (batch_size,), device=edge_index.device)
    return structured_negative_sampling(edge_index, rows=indices)
