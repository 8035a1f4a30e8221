from oracle.lightgcn_oracle import gcn_norm as _gcn_norm


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False, add_self_loops=True, flow="source_to_target",
             dtype=None):
    assert edge_weight is None and not add_self_loops and not improved, "stub covers the reference's call only"
    return _gcn_norm(edge_index, num_nodes)
