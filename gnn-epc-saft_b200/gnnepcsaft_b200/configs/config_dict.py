"""Attribute-style configuration container.

The reference builds `ml_collections.ConfigDict` objects (gnnepcsaft/configs/default.py:8) and only ever reads
them with attribute access (`config.hidden_dim`, train/utils.py:57-69; models.py:150-182) and overrides fields with
`--config.key=value`.  `ml_collections` is used when it is importable; otherwise this minimal stand-in gives the
same attribute / item / `to_dict` / `lock` surface, so `create_model(config, deg)` reads either."""

from __future__ import annotations


class _PlainConfigDict(dict):
    _locked = False

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        if key == "_locked":
            return object.__setattr__(self, key, value)
        if self._locked and key not in self:
            raise KeyError(f"config is locked; cannot add field '{key}'")
        self[key] = value

    def lock(self):
        self._locked = True
        return self

    def unlock(self):
        self._locked = False
        return self

    def to_dict(self):
        return dict(self)

    def update_from_flattened_dict(self, flat):
        for k, v in flat.items():
            setattr(self, k.split(".")[-1], v)


try:  # pragma: no cover - not installed in the build container
    from ml_collections import ConfigDict  # type: ignore
except ImportError:
    ConfigDict = _PlainConfigDict
