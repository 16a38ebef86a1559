"""CPU restatement of the reference's CIFAR score network (test infrastructure).

Restates, in PyTorch, the Flax module graph of

* cifar/models/ddpm.py:47-101          ScoreNet.__call__
* cifar/models/layers.py:95-107        ddpm_conv3x3  (nn.Conv 3x3, SAME, bias)
* cifar/models/layers.py:450-461       get_timestep_embedding
* cifar/models/layers.py:464-475       NIN
* cifar/models/layers.py:493-511       AttnBlock
* cifar/models/layers.py:514-537       Upsample / Downsample
* cifar/models/layers.py:540-565       ResnetBlockDDPM
* cifar/models/normalization.py:38-39  nn.GroupNorm (flax 0.9.0 defaults)
* cifar/models/layers.py:60-63         default_init

Parameters live in a nested dict with the names Flax's ``nn.compact``
auto-naming gives them (SURVEY.md §5 "Checkpoint / resume"), kernels in Flax
layout (conv HWIO, dense (in, out)), so a real ``params_ema`` pytree converted
leaf-by-leaf to torch tensors drops in.

Flax semantics restated from documentation (flax/jax are absent -> parity
unpinned, see oracle/__init__.py):
  - SAME padding with stride 2 on even H pads (0, 1)           (Appendix C.1)
  - GroupNorm: 32 groups, eps 1e-6, var = E[x^2] - E[x]^2 >= 0 (Appendix C.2)
  - raw continuous t feeds the sinusoidal embedding            (Appendix C.3)
  - jax.image.resize(..., 'nearest') x2 = index // 2           (Appendix C.9)
"""
import math

import torch
import torch.nn.functional as F


def swish(x):
    return x * torch.sigmoid(x)


def get_timestep_embedding(timesteps, embedding_dim, max_positions=10000):
    # cifar/models/layers.py:450-461
    assert timesteps.dim() == 1
    half_dim = embedding_dim // 2
    emb = math.log(max_positions) / (half_dim - 1)
    emb = torch.exp(torch.arange(half_dim, dtype=torch.float32) * -emb).to(timesteps.dtype)
    emb = timesteps[:, None] * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def group_norm(x, p, num_groups=32, eps=1e-6):
    # flax.linen.GroupNorm over an NHWC tensor: statistics over (H, W, C/G)
    B, H, W, C = x.shape
    g = x.reshape(B, H * W, num_groups, C // num_groups)
    mean = g.mean(dim=(1, 3), keepdim=True)
    mean2 = (g * g).mean(dim=(1, 3), keepdim=True)
    var = torch.clamp(mean2 - mean * mean, min=0.0)
    y = (g - mean) * torch.rsqrt(var + eps)
    y = y.reshape(B, H, W, C)
    return y * p["scale"] + p["bias"]


def conv3x3(x, p, stride=1):
    # nn.Conv(kernel (3,3), padding='SAME'), NHWC activations, HWIO kernel
    w = p["kernel"].permute(3, 2, 0, 1)  # OIHW
    xn = x.permute(0, 3, 1, 2)
    if stride == 1:
        y = F.conv2d(xn, w, p["bias"], stride=1, padding=1)
    else:
        xn = F.pad(xn, (0, 1, 0, 1))     # SAME, stride 2, even size -> (0, 1)
        y = F.conv2d(xn, w, p["bias"], stride=2, padding=0)
    return y.permute(0, 2, 3, 1)


def dense(x, p):
    return x @ p["kernel"] + p["bias"]


def nin(x, p):
    # cifar/models/layers.py:464-475
    return torch.tensordot(x, p["W"], dims=1) + p["b"]


def attn_block(x, p):
    # cifar/models/layers.py:493-511
    B, H, W, C = x.shape
    h = group_norm(x, p["GroupNorm_0"])
    q = nin(h, p["NIN_0"])
    k = nin(h, p["NIN_1"])
    v = nin(h, p["NIN_2"])
    w = torch.einsum("bhwc,bHWc->bhwHW", q, k) * (int(C) ** (-0.5))
    w = w.reshape(B, H, W, H * W)
    w = torch.softmax(w, dim=-1)
    w = w.reshape(B, H, W, H, W)
    h = torch.einsum("bhwHW,bHWc->bhwc", w, v)
    h = nin(h, p["NIN_3"])
    return x + h


def upsample(x, p, with_conv=True):
    # cifar/models/layers.py:514-523
    h = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    if with_conv:
        h = conv3x3(h, p["Conv_0"])
    return h


def downsample(x, p, with_conv=True):
    # cifar/models/layers.py:526-537
    if with_conv:
        return conv3x3(x, p["Conv_0"], stride=2)
    xn = x.permute(0, 3, 1, 2)
    return F.avg_pool2d(xn, 2, 2).permute(0, 2, 3, 1)


def resnet_block(x, temb, p, out_ch=None):
    # cifar/models/layers.py:540-565 (train=False -> dropout is the identity)
    C = x.shape[-1]
    out_ch = out_ch if out_ch else C
    h = swish(group_norm(x, p["GroupNorm_0"]))
    h = conv3x3(h, p["Conv_0"])
    if temb is not None:
        h = h + dense(swish(temb), p["Dense_0"])[:, None, None, :]
    h = swish(group_norm(h, p["GroupNorm_1"]))
    h = conv3x3(h, p["Conv_1"])
    if C != out_ch:
        x = nin(x, p["NIN_0"])
    return x + h


def scorenet_apply(params, config, t, x, y=None):
    """cifar/models/ddpm.py:47-101 with train=False.  t: (B,1,1,1) or (B,),
    x: (B,H,W,C) NHWC, y: (B,) int labels or None.  Returns (B,H,W,C)."""
    m = config.model
    nf = m.nf
    ch_mult = tuple(m.ch_mult)
    num_res_blocks = m.num_res_blocks
    attn_resolutions = tuple(m.attn_resolutions)
    num_resolutions = len(ch_mult)
    n_res = n_attn = n_down = n_up = 0

    temb = get_timestep_embedding(t.reshape(-1), nf)                 # :64
    temb = dense(temb, params["Dense_0"])                            # :65
    temb = dense(swish(temb), params["Dense_1"])                     # :66
    if m.conditioned:
        temb = temb + params["Embed_0"]["embedding"][y.long()]       # :68

    hs = [conv3x3(x, params["Conv_0"])]                              # :71
    for i_level in range(num_resolutions):
        for _ in range(num_res_blocks):
            h = resnet_block(hs[-1], temb, params[f"ResnetBlockDDPM_{n_res}"],
                             out_ch=nf * ch_mult[i_level])           # :75
            n_res += 1
            if h.shape[1] in attn_resolutions:
                h = attn_block(h, params[f"AttnBlock_{n_attn}"])     # :77
                n_attn += 1
            hs.append(h)
        if i_level != num_resolutions - 1:
            hs.append(downsample(hs[-1], params[f"Downsample_{n_down}"], m.resamp_with_conv))  # :80
            n_down += 1

    h = hs[-1]
    h = resnet_block(h, temb, params[f"ResnetBlockDDPM_{n_res}"]); n_res += 1   # :83
    h = attn_block(h, params[f"AttnBlock_{n_attn}"]); n_attn += 1               # :84
    h = resnet_block(h, temb, params[f"ResnetBlockDDPM_{n_res}"]); n_res += 1   # :85

    for i_level in reversed(range(num_resolutions)):
        for _ in range(num_res_blocks + 1):
            h = resnet_block(torch.cat([h, hs.pop()], dim=-1), temb,
                             params[f"ResnetBlockDDPM_{n_res}"], out_ch=nf * ch_mult[i_level])  # :90
            n_res += 1
        if h.shape[1] in attn_resolutions:
            h = attn_block(h, params[f"AttnBlock_{n_attn}"])         # :92
            n_attn += 1
        if i_level != 0:
            h = upsample(h, params[f"Upsample_{n_up}"], m.resamp_with_conv)  # :94
            n_up += 1
    assert not hs                                                    # :96
    h = swish(group_norm(h, params["GroupNorm_0"]))                  # :98
    h = conv3x3(h, params["Conv_1"])                                 # :99
    return h


def params_to(params, dtype=None, device=None):
    if isinstance(params, dict):
        return {k: params_to(v, dtype, device) for k, v in params.items()}
    return params.to(dtype=dtype, device=device)


# ---------------------------------------------------------------------------
# Toy MLP of the notebook (notebooks/superposition_edu.ipynb:157-173)
# ---------------------------------------------------------------------------

def toy_mlp_apply(params, t, x):
    """hstack([t, x]) -> 4 x (Dense 512 + swish) -> Dense ndim."""
    h = torch.cat([t, x], dim=1)
    for i in range(4):
        h = swish(dense(h, params[f"Dense_{i}"]))
    return dense(h, params["Dense_4"])
