"""reference gnnepcsaft/configs/default.py."""
from gnnepcsaft_b200.configs.default import get_config  # noqa: F401
