"""Default hyper-parameters: field-for-field the values of reference gnnepcsaft/configs/default.py:6-46."""

from .config_dict import ConfigDict


def get_config():
    """Get the default hyperparameter configuration."""
    config = ConfigDict()
    config.job_type = "train"
    # optimizer (reference default.py:11-15)
    config.optimizer = "adam"
    config.learning_rate = 1.0e-3
    config.change_opt = False
    config.change_sch = False
    # training hyper-parameters (default.py:19-33)
    config.accelerator = "gpu"
    config.batch_size = 512
    config.pad_size = 128
    config.num_train_steps = 250_000
    config.warmup_steps = 100
    config.log_every_steps = 10_000
    config.eval_every_steps = 24_999
    config.checkpoint_every_steps = 25000
    config.amp = False
    config.momentum = 0.9
    config.patience = 5
    config.weight_decay = 1e-2
    config.dataset = "esper"
    config.checkpoint = ""
    # GNN hyper-parameters (default.py:36-46)
    config.model_name = "esper1"
    config.model = "PNAL"
    config.propagation_depth = 6
    config.hidden_dim = 64
    config.pre_layers = 1
    config.post_layers = 1
    config.dropout_rate = 0.0
    config.num_mlp_layers = 1
    config.num_para = 5
    config.skip_connections = True
    config.add_self_loops = True
    return config
