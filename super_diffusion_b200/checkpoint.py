"""Parameter-tree import / export for the score-net.

The reference keeps its models in orbax checkpoints of a ``State`` whose ``params_ema`` pytree uses Flax's
auto-generated module names and layouts (cifar/run_lib.py:43-52, cifar/models/utils.py:30-39; names pinned against the
reference's own init in tests/test_reference_vectors.py).  orbax / tensorstore are not available here, so the importer
takes the two portable dumps a reference user can produce in three lines next to their checkpoint::

    state = ckpt_mgr.restore(step, items=state)                                  # cifar/run_lib.py:50-52
    np.savez("modelA.npz", **flatten_dict(state.params_ema, sep="/"))            # (a) flat .npz, keys "Conv_0/kernel"
    open("modelA.msgpack", "wb").write(flax.serialization.to_bytes(state.params_ema))   # (b) Flax msgpack

(b) is decoded here without Flax: ``flax.serialization`` writes a msgpack map whose array leaves are
``ExtType(1, packb((shape, dtype_name, raw_bytes)))`` (format restated from the Flax source; no reference test pins it).
Both loaders return the nested dict of torch tensors that ``ScoreNet.bind`` / ``mutils.State(params_ema=...)`` take, and
``validate_params`` checks names and shapes against the configuration before anything reaches the GPU.
"""
import numpy as np
import torch

from .models import utils as mutils

_EXT_NDARRAY = 1


def flatten_params(params, prefix="", sep="/"):
    out = {}
    for k, v in params.items():
        key = f"{prefix}{sep}{k}" if prefix else str(k)
        if isinstance(v, dict):
            out.update(flatten_params(v, key, sep))
        else:
            out[key] = v
    return out


def unflatten_params(flat, sep="/"):
    out = {}
    for key, v in flat.items():
        node = out
        parts = key.split(sep)
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise ValueError(f"key {key!r} nests under a leaf")
        if parts[-1] in node:
            raise ValueError(f"duplicate key {key!r}")
        node[parts[-1]] = v
    return out


def _to_tensor(a):
    a = np.asarray(a)
    if a.dtype.kind not in "fiu":
        raise ValueError(f"unsupported leaf dtype {a.dtype}")
    t = torch.from_numpy(np.array(a, copy=True, order="C"))      # frombuffer views are read-only
    return t.float() if a.dtype.kind == "f" else t


def save_npz(path, params):
    np.savez(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                      for k, v in flatten_params(params).items()})


def load_npz(path):
    with np.load(path) as z:
        return unflatten_params({k: _to_tensor(z[k]) for k in z.files})


def to_msgpack_bytes(params):
    """Encode like flax.serialization.to_bytes (array leaves -> ExtType 1)."""
    import msgpack

    def enc(node):
        if isinstance(node, dict):
            return {k: enc(v) for k, v in node.items()}
        a = node.detach().cpu().numpy() if torch.is_tensor(node) else np.asarray(node)
        return msgpack.ExtType(_EXT_NDARRAY, msgpack.packb((list(a.shape), a.dtype.name, a.tobytes("C")), use_bin_type=True))
    return msgpack.packb(enc(params), use_bin_type=True)


def from_msgpack_bytes(data):
    import msgpack

    def ext_hook(code, payload):
        if code != _EXT_NDARRAY:
            raise ValueError(f"unsupported Flax msgpack extension type {code}")
        shape, dtype_name, buf = msgpack.unpackb(payload, raw=True)
        name = dtype_name.decode() if isinstance(dtype_name, bytes) else dtype_name
        if name == "bfloat16":
            raw = np.frombuffer(buf, dtype=np.uint16).astype(np.uint32) << 16
            arr = raw.view(np.float32)
        else:
            arr = np.frombuffer(buf, dtype=np.dtype(name))
        return _to_tensor(arr.reshape([int(s) for s in shape]))

    tree = msgpack.unpackb(data, ext_hook=ext_hook, raw=False, strict_map_key=False)
    if not isinstance(tree, dict):
        raise ValueError("msgpack checkpoint does not hold a parameter dict")
    return tree.get("params", tree) if set(tree.keys()) == {"params"} else tree


def load_msgpack(path):
    with open(path, "rb") as fh:
        return from_msgpack_bytes(fh.read())


def load_params(path):
    """Dispatch on the file extension: .npz (flat keys joined with '/') or .msgpack (Flax)."""
    p = str(path)
    if p.endswith(".npz"):
        return load_npz(p)
    if p.endswith(".msgpack") or p.endswith(".flax"):
        return load_msgpack(p)
    raise ValueError(f"unknown checkpoint format: {p} (expected .npz or .msgpack; orbax directories must be exported "
                     "with the three-line recipe in super_diffusion_b200/checkpoint.py)")


def expected_shapes(config):
    """Names and shapes of the 'score-net' parameter tree for ``config`` (cifar/models/ddpm.py:47-101)."""
    ref = mutils.init_scorenet_params(torch.Generator().manual_seed(0), config)
    return {k: tuple(v.shape) for k, v in flatten_params(ref).items()}


def validate_params(params, config):
    """Raise ValueError naming every missing / unexpected / mis-shaped leaf; returns the tree unchanged."""
    want = expected_shapes(config)
    got = {k: tuple(v.shape) for k, v in flatten_params(params).items()}
    missing = sorted(set(want) - set(got))
    extra = sorted(set(got) - set(want))
    wrong = sorted(k for k in set(want) & set(got) if want[k] != got[k])
    if missing or extra or wrong:
        msg = []
        if missing:
            msg.append(f"missing {missing[:6]}{'...' if len(missing) > 6 else ''}")
        if extra:
            msg.append(f"unexpected {extra[:6]}{'...' if len(extra) > 6 else ''}")
        if wrong:
            msg.append("shape mismatch " + ", ".join(f"{k}: {got[k]} != {want[k]}" for k in wrong[:6]))
        raise ValueError("parameter tree does not match config.model: " + "; ".join(msg))
    return params


def restore_state(path, config):
    """Checkpoint file -> ``State`` with ``params_ema`` (what the joint vector fields read, cifar/dynamics.py:113)."""
    params = validate_params(load_params(path), config)
    return mutils.State(step=0, model_params=params, params_ema=params, ema_rate=config.model.ema_rate)
