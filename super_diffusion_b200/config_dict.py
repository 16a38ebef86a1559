"""Minimal stand-in for ``ml_collections.ConfigDict`` (absent from this image).

The reference builds its configs with ml_collections
(cifar/configs/sm/cifar/vpsde.py:1-60) and reads them by attribute
(cifar/models/ddpm.py:49-58, cifar/eval_utils.py:48-51).  Only attribute /
item access, nesting, ``lock()`` and ``to_dict()`` are provided.
"""


class ConfigDict(dict):
    def __init__(self, initial=None):
        super().__init__()
        object.__setattr__(self, "_locked", False)
        if initial:
            for k, v in initial.items():
                self[k] = ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        if self._locked and name not in self:
            raise AttributeError(f"config is locked; cannot add field {name!r}")
        self[name] = value

    def lock(self):
        object.__setattr__(self, "_locked", True)
        for v in self.values():
            if isinstance(v, ConfigDict):
                v.lock()
        return self

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self.items()}
