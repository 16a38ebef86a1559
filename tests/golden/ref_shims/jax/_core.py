import numpy as np


class _Config:
    x64 = False

    def update(self, name, value):
        if name == "jax_enable_x64":
            self.x64 = bool(value)
        else:
            raise KeyError(name)


config = _Config()


def default_float():
    return np.float64 if config.x64 else np.float32


class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, value):
        out = np.array(self.arr, copy=True).view(JArr)
        out[self.idx] = value
        return out

    def add(self, value):
        out = np.array(self.arr, copy=True).view(JArr)
        out[self.idx] += value
        return out


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


class JArr(np.ndarray):
    """ndarray with the functional `.at[idx].set(v)` update of jax arrays."""

    @property
    def at(self):
        return _At(self)


def asjarr(x):
    if isinstance(x, np.ndarray):
        return x.view(JArr)
    if isinstance(x, np.generic):
        # 0-d results behave like JAX weakly-typed scalars: they must not upcast float32 arrays
        return x.item()
    return x
