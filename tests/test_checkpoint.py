"""Parameter-tree import and the CLI's host logic -- CPU only."""
import numpy as np
import pytest
import torch

from super_diffusion_b200 import checkpoint as ckpt
from super_diffusion_b200 import main as cli
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import ddpm  # noqa: F401  (registers 'score-net')
from super_diffusion_b200.models import utils as mutils


def _small_config(conditioned=False):
    cfg = vpsde.get_config(conditioned=conditioned)
    cfg.model.nf = 64
    cfg.model.ch_mult = (1, 2)
    cfg.model.num_res_blocks = 1
    cfg.model.attn_resolutions = (16,)
    return cfg


def _equal(a, b):
    fa, fb = ckpt.flatten_params(a), ckpt.flatten_params(b)
    assert fa.keys() == fb.keys()
    for k in fa:
        assert torch.equal(fa[k], fb[k]), k


def test_npz_and_msgpack_round_trip(tmp_path):
    cfg = _small_config(conditioned=True)
    _, params = mutils.init_model(3, cfg)
    ckpt.save_npz(tmp_path / "a.npz", params)
    _equal(ckpt.load_params(tmp_path / "a.npz"), params)
    (tmp_path / "a.msgpack").write_bytes(ckpt.to_msgpack_bytes(params))
    _equal(ckpt.load_params(tmp_path / "a.msgpack"), params)
    # a {'params': tree} wrapper (what model.init returns) is unwrapped
    _equal(ckpt.from_msgpack_bytes(ckpt.to_msgpack_bytes({"params": params})), params)
    st = ckpt.restore_state(tmp_path / "a.npz", cfg)
    assert st.params_ema is st.model_params and st.ema_rate == cfg.model.ema_rate
    with pytest.raises(ValueError):
        ckpt.load_params(tmp_path / "a.ckpt")


def test_flax_msgpack_wire_format():
    """Hand-built bytes in flax.serialization's layout: map -> ExtType(1, packb((shape, dtype name, C-order bytes)))."""
    import msgpack
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.array([1.5, -2.0], dtype=np.float64)
    h = (np.array([1.0, -0.5, 3.0], dtype=np.float32).view(np.uint32) >> 16).astype(np.uint16)      # bfloat16 payload
    ext = lambda shape, name, raw: msgpack.ExtType(1, msgpack.packb((shape, name, raw), use_bin_type=True))
    blob = msgpack.packb({"Dense_0": {"kernel": ext((2, 3), "float32", a.tobytes()), "bias": ext((2,), "float64", b.tobytes())},
                          "half": ext((3,), "bfloat16", h.tobytes())}, use_bin_type=True)
    tree = ckpt.from_msgpack_bytes(blob)
    assert torch.equal(tree["Dense_0"]["kernel"], torch.from_numpy(a))
    assert tree["Dense_0"]["bias"].dtype == torch.float32 and tree["Dense_0"]["bias"].tolist() == [1.5, -2.0]
    assert tree["half"].tolist() == [1.0, -0.5, 3.0]
    with pytest.raises(ValueError):
        ckpt.from_msgpack_bytes(msgpack.packb({"x": msgpack.ExtType(2, b"")}))


def test_validate_params_names_every_problem():
    cfg = _small_config()
    _, params = mutils.init_model(0, cfg)
    assert ckpt.validate_params(params, cfg) is params
    flat = ckpt.flatten_params(params)
    assert len(flat) == len(ckpt.expected_shapes(cfg))
    bad = dict(flat)
    del bad["Conv_0/kernel"]
    bad["Conv_9/kernel"] = torch.zeros(1)
    bad["Dense_0/kernel"] = torch.zeros(3, 3)
    with pytest.raises(ValueError) as e:
        ckpt.validate_params(ckpt.unflatten_params(bad), cfg)
    msg = str(e.value)
    assert "Conv_0/kernel" in msg and "Conv_9/kernel" in msg and "Dense_0/kernel" in msg
    # a conditioned checkpoint does not fit an unconditioned config (extra Embed_0)
    _, pc = mutils.init_model(0, _small_config(conditioned=True))
    with pytest.raises(ValueError):
        ckpt.validate_params(pc, cfg)
    with pytest.raises(ValueError):
        ckpt.unflatten_params({"a": torch.zeros(1), "a/b": torch.zeros(1)})


def test_cli_config_and_modes(tmp_path):
    assert cli.load_config("vpsde").model.conditioned is False
    assert cli.load_config("vpsdeA").model.conditioned is True and cli.load_config("vpsdeB").data.train_split == "train[50%:]"
    f = tmp_path / "cfg.py"
    f.write_text("from super_diffusion_b200.configs import vpsde\n\ndef get_config():\n    c = vpsde.get_config()\n    c.seed = 7\n    return c\n")
    assert cli.load_config(str(f)).seed == 7
    with pytest.raises(ValueError):
        cli.load_config("nope")
    args = cli.build_parser().parse_args(["--config", "vpsde", "--workdir", "w", "--mode", "eval_joint_fid_stoch", "--chkpts", "a.npz, b.npz"])
    assert args.mode == "eval_joint_fid_stoch" and args.eval_folder == "eval"
    with pytest.raises(SystemExit):
        cli.launch(["--config", "vpsde", "--workdir", "w", "--mode", "train"])
    with pytest.raises(SystemExit):
        cli.launch(["--config", "vpsde", "--workdir", "w", "--mode", "eval_joint_fid_stoch"])


def _write_zarr_leaf(dirpath, arr, chunks=None, compressor=None):
    """A leaf in tensorstore's zarr v2 layout (what orbax's PyTreeCheckpointHandler writes per array without OCDBT)."""
    import gzip
    import json
    import os
    os.makedirs(dirpath)
    chunks = list(chunks or arr.shape)
    meta = {"zarr_format": 2, "shape": list(arr.shape), "chunks": chunks, "dtype": arr.dtype.str, "order": "C", "fill_value": 0,
            "filters": None, "compressor": {"id": compressor} if compressor else None, "dimension_separator": "."}
    with open(os.path.join(dirpath, ".zarray"), "w") as fh:
        json.dump(meta, fh)
    grid = [-(-s // c) for s, c in zip(arr.shape, chunks)]
    for idx in np.ndindex(*grid):
        block = np.zeros(chunks, dtype=arr.dtype)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, arr.shape))
        block[tuple(slice(0, s.stop - s.start) for s in sl)] = arr[sl]
        raw = block.tobytes("C")
        with open(os.path.join(dirpath, ".".join(str(i) for i in idx)), "wb") as fh:
            fh.write(gzip.compress(raw) if compressor == "gzip" else raw)


def test_orbax_directory_zarr_layout(tmp_path):
    """<workdir>/checkpoints/chkpt_<step>/default/<item>.<dotted key>/ (cifar/run_lib.py:43-52): the latest step is picked,
    params_ema is read leaf by leaf (chunked, gzip or raw), other State items are ignored."""
    cfg = vpsde.get_config()
    cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks, cfg.model.attn_resolutions = 64, (1, 2), 1, (16,)   # 1.5 M parameters
    _, params = mutils.init_model(3, cfg, zero_init_scale=1.0)
    flat = ckpt.flatten_params(params, sep=".")
    root = tmp_path / "checkpoints"
    for step, scale in ((5000, 0.5), (10000, 1.0)):
        d = root / f"chkpt_{step}" / "default"
        for i, (k, v) in enumerate(flat.items()):
            a = (v * scale).numpy()
            chunks = [max(1, s // 2 + 1) for s in a.shape] if i % 3 == 0 else None
            _write_zarr_leaf(str(d / f"params_ema.{k}"), a, chunks=chunks, compressor="gzip" if i % 2 else None)
        _write_zarr_leaf(str(d / "model_params.Conv_0.bias"), np.ones(64, dtype=np.float32))
    got = ckpt.validate_params(ckpt.load_params(str(root)), cfg)
    for k, v in ckpt.flatten_params(got, sep=".").items():
        assert torch.equal(v, flat[k]), k
    half = ckpt.load_orbax(str(root / "chkpt_5000"))
    assert torch.equal(half["Conv_0"]["kernel"], params["Conv_0"]["kernel"] * 0.5)
    (root / "chkpt_20000" / "default").mkdir(parents=True)
    (root / "chkpt_20000" / "default" / "manifest.ocdbt").write_bytes(b"")
    with pytest.raises(RuntimeError, match="OCDBT"):
        ckpt.load_params(str(root))


def test_orbax_directory_through_orbax_when_importable(tmp_path, monkeypatch):
    """With orbax importable (a reference user's environment) the directory is restored by orbax itself; exercised here
    against a stub module, since orbax is absent from the build image."""
    import sys
    import types
    cfg = vpsde.get_config()
    cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks, cfg.model.attn_resolutions = 64, (1, 2), 1, (16,)
    _, params = mutils.init_model(4, cfg, zero_init_scale=1.0)
    seen = {}

    class PyTreeCheckpointer:
        def restore(self, d):
            seen["dir"] = d
            to_np = lambda t: {k: to_np(v) for k, v in t.items()} if isinstance(t, dict) else t.numpy()
            return {"step": 7, "params_ema": to_np(params), "model_params": {}}
    ocp = types.ModuleType("orbax.checkpoint")
    ocp.PyTreeCheckpointer = PyTreeCheckpointer
    pkg = types.ModuleType("orbax")
    pkg.checkpoint = ocp
    monkeypatch.setitem(sys.modules, "orbax", pkg)
    monkeypatch.setitem(sys.modules, "orbax.checkpoint", ocp)
    d = tmp_path / "checkpoints" / "chkpt_12" / "default"
    d.mkdir(parents=True)
    got = ckpt.validate_params(ckpt.load_params(str(tmp_path / "checkpoints")), cfg)
    assert seen["dir"] == str(d)
    assert torch.equal(got["Conv_1"]["kernel"], params["Conv_1"]["kernel"])
