import dataclasses


def dataclass(cls):
    return dataclasses.dataclass(cls)
