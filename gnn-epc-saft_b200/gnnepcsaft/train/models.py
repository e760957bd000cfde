"""reference gnnepcsaft/train/models.py surface on the sm_100a kernels."""
from gnnepcsaft_b200.train.models import (PNAPCSAFT, PNApcsaftL, PnaconvsParams,  # noqa: F401
                                          ReadoutMLPParams, mape)

__b200_shim__ = True
