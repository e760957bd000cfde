"""reference gnnepcsaft/train/utils.py surface (hot-path part) on the sm_100a kernels."""
from gnnepcsaft_b200.train.utils import (calc_deg, create_model, create_optimizer,  # noqa: F401
                                         load_checkpoint, savemodel)

__b200_shim__ = True
