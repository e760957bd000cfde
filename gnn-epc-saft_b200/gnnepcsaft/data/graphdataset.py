"""reference gnnepcsaft/data/graphdataset.py readers: the processed `(Data, slices)` files as packed stores."""
from gnnepcsaft_b200.data.dataset_io import Esper, Ramirez, load_inmemory_dataset  # noqa: F401
