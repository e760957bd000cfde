"""The six architectures of reference gnnepcsaft/configs/configs_parallel.py:4-62 (entry 4 is BASELINE config 5's
wide/deep variant, entry 3 the pre_layers = 2 model that needs the general per-edge message path)."""

_DEFAULT = dict(propagation_depth=6, hidden_dim=64, num_mlp_layers=1, pre_layers=1, post_layers=1,
                skip_connections=True, add_self_loops=True)


def get_configs() -> list:
    """Get the hyperparameter configurations."""
    return [
        dict(_DEFAULT),
        dict(_DEFAULT),
        dict(_DEFAULT),
        dict(_DEFAULT, propagation_depth=3, num_mlp_layers=0, pre_layers=2),
        dict(_DEFAULT, hidden_dim=256, post_layers=2, add_self_loops=False),
        dict(_DEFAULT),
    ]
