"""CIFAR-10 VP-SDE score-net config; same fields and values as the reference's
cifar/configs/sm/cifar/vpsde.py:4-60 (variants vpsdeA/B, vpsde_less_5/more_5
differ only in data.train_split and model.conditioned)."""
from ..config_dict import ConfigDict


def get_config(conditioned=False, train_split="train"):
    config = ConfigDict()
    config.seed = 1

    config.data = data = ConfigDict()
    data.dataset = "CIFAR10"
    data.train_split = train_split
    data.ndims = 3
    data.image_size = 32
    data.num_channels = 3
    data.num_classes = 10
    data.uniform_dequantization = True
    data.random_flip = True
    data.task = "generate"
    data.dynamics = "vpsde"
    data.t_0, data.t_1 = 0.0, 1.0

    config.model = model = ConfigDict()
    model.name = "score-net"
    model.conditioned = conditioned
    model.loss = "dsm"
    model.ema_rate = 0.9999
    model.normalization = "GroupNorm"
    model.nonlinearity = "swish"
    model.nf = 128
    model.ch_mult = (1, 2, 2, 2)
    model.num_res_blocks = 2
    model.attn_resolutions = (16, 8)
    model.resamp_with_conv = True
    model.dropout = 0.1

    config.train = train = ConfigDict()
    train.batch_size = 128
    train.n_jitted_steps = 1
    train.n_iters = 500_000
    train.save_every = 5_000
    train.eval_every = 10_000
    train.log_every = 50
    train.lr = 2e-4
    train.beta1 = 0.9
    train.eps = 1e-8
    train.warmup = 5_000
    train.grad_clip = 1.0

    config.eval = ev = ConfigDict()
    ev.batch_size = 100
    ev.artifact_size = 64
    ev.num_samples = 50_000
    ev.use_ema = True
    ev.estimate_bpd = True
    return config


def get_config_A():
    """cifar/configs/sm/cifar/vpsdeA.py: first half of train, class-conditioned."""
    return get_config(conditioned=True, train_split="train[:50%]")


def get_config_B():
    """cifar/configs/sm/cifar/vpsdeB.py: second half of train, class-conditioned."""
    return get_config(conditioned=True, train_split="train[50%:]")
