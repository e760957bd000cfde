"""B200-native (sm_100a) implementation of the GNN-ePC-SAFT PNAPCSAFT message-passing path."""

from .data.graph import Batch, Data, collate  # noqa: F401

__version__ = "0.1.0"
