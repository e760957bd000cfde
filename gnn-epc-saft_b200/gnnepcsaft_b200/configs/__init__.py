"""Configuration objects of the reference (gnnepcsaft/configs/default.py, configs_parallel.py)."""
from .config_dict import ConfigDict  # noqa: F401
