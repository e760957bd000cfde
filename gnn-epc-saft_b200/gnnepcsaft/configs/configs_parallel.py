"""reference gnnepcsaft/configs/configs_parallel.py."""
from gnnepcsaft_b200.configs.configs_parallel import get_configs  # noqa: F401
