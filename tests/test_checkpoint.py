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
