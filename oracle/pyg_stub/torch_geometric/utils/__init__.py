from oracle.lightgcn_oracle import structured_negative_sampling  # noqa: F401
