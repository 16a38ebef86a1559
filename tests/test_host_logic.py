"""CPU tests of the host-side mirror of the reference API (no GPU needed)."""
import numpy as np
import pytest
import torch

from oracle import schedule as OSched
from oracle import scorenet as OS
from super_diffusion_b200 import sde
from super_diffusion_b200.config_dict import ConfigDict
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import utils as mutils
from super_diffusion_b200 import distributed as D


def test_config_matches_reference_fields():
    cfg = vpsde.get_config()
    assert cfg.model.name == "score-net" and cfg.model.nf == 128 and tuple(cfg.model.ch_mult) == (1, 2, 2, 2)
    assert cfg.model.num_res_blocks == 2 and tuple(cfg.model.attn_resolutions) == (16, 8)
    assert cfg.data.image_size == 32 and cfg.data.num_channels == 3 and cfg.eval.batch_size == 100
    assert not cfg.model.conditioned and vpsde.get_config_A().model.conditioned
    assert vpsde.get_config_B().data.train_split == "train[50%:]"
    cfg.lock()
    with pytest.raises(AttributeError):
        cfg.model.new_field = 1
    assert isinstance(cfg.to_dict()["model"], dict)


def test_registry_and_param_tree():
    assert mutils.get_model("score-net").__name__ == "ScoreNet"
    with pytest.raises(ValueError):
        mutils.register_model(name="score-net")(type("X", (), {}))
    cfg = vpsde.get_config()
    _, params = mutils.init_model(0, cfg)
    n = mutils.count_params(params)
    assert abs(n - 36.01e6) < 0.05e6, n          # 36.01 M parameters
    assert len([k for k in params if k.startswith("ResnetBlockDDPM_")]) == 22
    assert len([k for k in params if k.startswith("AttnBlock_")]) == 7
    assert params["Conv_0"]["kernel"].shape == (3, 3, 3, 128) and params["Conv_1"]["kernel"].shape == (3, 3, 128, 3)
    assert "NIN_0" in params["ResnetBlockDDPM_2"] and "NIN_0" not in params["ResnetBlockDDPM_0"]
    _, pc = mutils.init_model(0, vpsde.get_config(conditioned=True))
    assert pc["Embed_0"]["embedding"].shape == (10, 512)


def test_oracle_scorenet_runs_and_faithful_init_is_near_zero():
    cfg = ConfigDict(vpsde.get_config().to_dict())
    cfg.model.nf = 64
    cfg.model.ch_mult = (1, 2)
    cfg.model.attn_resolutions = (8,)
    cfg.data.image_size = 16
    _, params = mutils.init_model(0, cfg)
    x = torch.randn(2, 16, 16, 3)
    out = OS.scorenet_apply(params, cfg, torch.full((2, 1, 1, 1), 0.5), x, None)
    assert out.shape == x.shape and out.abs().max() < 1e-3
    _, p2 = mutils.init_model(0, cfg, zero_init_scale=1.0)
    out2 = OS.scorenet_apply(p2, cfg, torch.full((2, 1, 1, 1), 0.5), x, None)
    assert out2.abs().max() > 1e-2
    # batch independence of the oracle itself
    out3 = OS.scorenet_apply(p2, cfg, torch.full((1, 1, 1, 1), 0.5), x[1:], None)
    assert torch.allclose(out2[1:], out3, atol=1e-5)


def test_host_schedule_equals_oracle_schedule():
    for acc in ("float64", "float32"):
        assert np.array_equal(sde.time_grid(1000, 1e-3, acc), OSched.time_grid(1000, 1e-3, acc))
    ts = sde.time_grid(200, 5e-3)
    tab = sde.schedule_table(ts, 5e-3)
    assert tab.shape == (200, 4) and tab.dtype == torch.float32
    assert np.allclose(tab[:, 0].numpy(), OSched.dlog_alphadt(ts), rtol=1e-6)
    assert np.allclose(tab[:, 1].numpy(), OSched.beta(ts), rtol=1e-6)
    assert np.allclose(tab[:, 2].numpy(), ts, rtol=1e-6)
    a, b, c = sde.edm_sigmas(50)
    a2, b2, c2 = OSched.edm_sigmas(50)
    assert np.array_equal(a, a2) and np.array_equal(b, b2) and c == c2


def test_shard_bounds_cover_and_are_ragged_safe():
    for total in (0, 1, 7, 512, 8192, 16384 + 3):
        for world in (1, 2, 4, 8):
            bounds = [D.shard_bounds(total, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == total
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from super_diffusion_b200 import _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_device()
    cfg = vpsde.get_config()
    model, params = mutils.init_model(0, cfg)
    with pytest.raises(Exception):
        mutils.get_model_fn(model, params)(torch.zeros(1, 1, 1, 1), torch.zeros(1, 32, 32, 3), None)


def test_uint8_conversion_matches_reference_formula():
    """cifar/run_lib.py:244-245 + cifar/datasets.py:32-35: x*0.5+0.5, clip(x*255, 0, 255), truncating cast."""
    from super_diffusion_b200 import run_lib
    cfg = vpsde.get_config()
    x = torch.tensor([-3.0, -1.0, -0.5, 0.0, 0.5, 0.999, 1.0, 7.0])
    out = run_lib.to_uint8(x, cfg)
    ref = np.clip((x.numpy() * 0.5 + 0.5) * 255.0, 0.0, 255.0).astype(np.uint8)
    assert np.array_equal(out.numpy(), ref)
    assert run_lib.get_image_scaler(cfg)(torch.tensor(0.75)).item() == 0.5


def test_toy_mlp_module_is_the_notebook_mlp():
    """models/toy_mlp.py (superposition_edu.ipynb:157-173) against the oracle restatement on the same Flax-shaped tree."""
    import torch
    from oracle import scorenet as OS
    from super_diffusion_b200.models import utils as mutils
    from super_diffusion_b200.models.toy_mlp import MLP, get_sscore
    m = MLP.init(5)
    tree = m.to_flax()
    assert [tuple(tree[f"Dense_{i}"]["kernel"].shape) for i in range(5)] == [(3, 512), (512, 512), (512, 512), (512, 512), (512, 2)]
    assert sum(p.numel() for p in m.parameters()) == 791_042          # 0.79 M parameters
    g = torch.Generator().manual_seed(0)
    t, x = torch.rand(33, 1, generator=g), torch.randn(33, 2, generator=g)
    ref = OS.toy_mlp_apply(OS.params_to(tree, dtype=torch.float64), t.double(), x.double())
    assert (get_sscore(m)(t, x).double() - ref).abs().max() < 1e-5
    assert torch.equal(MLP.from_flax({"params": tree})(t, x), m(t, x))
    assert mutils.get_model("toy-mlp") is MLP
