from . import train_state  # noqa: F401
