class ConfigDict(dict):
    """Attribute-access dict, enough for cifar/configs/sm/cifar/*.py."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v
