class TrainState:
    def __init__(self, apply_fn=None, params=None, tx=None):
        self.apply_fn, self.params, self.tx = apply_fn, params, tx

    @classmethod
    def create(cls, *, apply_fn, params, tx=None):
        return cls(apply_fn, params, tx)
