"""Sample-generation driver — the sampling half of the reference's cifar/run_lib.py::evaluate_joint_fid
(:201-259) and evaluate_fid (:129-167): batch loop, vector-field choice, inverse scaler, uint8 conversion and the
``samples_{batch_id}.npz`` on-disk format (keys ``samples``, ``num_steps``).  SURVEY.md §8(f) row N2.

The Inception / FID half (:262-278, cifar/evaluation.py, TensorFlow) is a quality metric, not part of the sampling
path, and is not reproduced; the ``samples_*.npz`` files written here are what the reference's own
``statistics`` / FID stage consumes.
"""
import math
import os

import numpy as np
import torch

from . import checkpoint, dynamics, eval_utils
from . import distributed as dist_utils
from .models import utils as mutils


def get_image_scaler(config):
    """cifar/datasets.py:26-29."""
    return lambda x: (x - 0.5) / 0.5


def get_image_inverse_scaler(config):
    """cifar/datasets.py:32-35."""
    return lambda x: x * 0.5 + 0.5


def to_uint8(artifacts, config):
    """cifar/run_lib.py:244-245: inverse scaler, clip(x*255, 0, 255), truncating uint8 cast (jnp astype truncates)."""
    x = get_image_inverse_scaler(config)(artifacts)
    return torch.clamp(x * 255.0, 0.0, 255.0).to(torch.uint8)


def evaluate_joint_samples(config, workdir, eval_folder, params_list, stoch=True, num_batches=None, dt=None,
                           device=None, mode="or"):
    """Generate ``config.eval.num_samples`` SuperDiff samples from the models whose parameter trees are in
    ``params_list`` (the reference restores them from orbax checkpoints, :207-210) and write
    ``<workdir>/<eval_folder>/samples_stoch/samples_{i}.npz``.  Returns the sample directory."""
    if not stoch and mode == "and":
        raise NotImplementedError("SuperDiff-AND is defined by the noise of the stochastic step (superposition_edu.ipynb:899-905)")
    models, states = [], []
    for params in params_list:
        model = mutils.get_model(config.model.name)(config=config)
        models.append(model)
        states.append(mutils.State(params_ema=params, model_params=params))
    sample_dir = os.path.join(workdir, eval_folder, "samples_stoch" if stoch else "samples")      # :214-217
    os.makedirs(sample_dir, exist_ok=True)
    key = int(config.seed)
    if mode == "or" and not stoch:
        vector_field = dynamics.get_joint_vf(key, models, states)                 # :224-225 (ODE + Hutchinson divergence)
    elif mode == "or":
        vector_field = dynamics.get_joint_stoch_vf(key, models, states)           # :222-223
    elif mode == "and":
        vector_field = dynamics.get_joint_and_vf(key, models, states)
    else:
        vector_field = dynamics.get_avg_vf(key, models, states, stoch=stoch)      # evaluate_fid, :145
    generator = eval_utils.get_generator(models, config, vector_field, dt=dt, device=device)   # :226
    total = math.ceil(config.eval.num_samples / config.eval.batch_size)           # :238
    if num_batches is not None:
        total = min(total, num_batches)
    B = config.eval.batch_size // eval_utils.local_device_count()
    rank = dist_utils.rank()
    for batch_id in range(total):
        # one process per GPU: every rank generates its B-sample shard (the reference's vmap over local devices, :227,
        # :240-243), the shards are gathered over NCCL and rank 0 writes the batch
        labels = ((torch.arange(B) + rank * B) % config.data.num_classes).to(torch.int32)      # tile(arange(10), 10), :242
        artifacts, num_steps = generator((key * 100003 + batch_id + 1) * 64 + rank, labels)
        artifacts = dist_utils.gather_samples(artifacts, B * dist_utils.world_size())
        if rank == 0:
            arr = to_uint8(artifacts, config).cpu().numpy()
            with open(os.path.join(sample_dir, f"samples_{batch_id}.npz"), "wb") as fout:
                np.savez_compressed(fout, samples=arr, num_steps=num_steps)        # :248-251
    return sample_dir


def evaluate_joint_fid(config, workdir, eval_folder, checkpoints, stoch, num_batches=None, dt=None, device=None):
    """cifar/run_lib.py:201-278 up to the Inception stage: SuperDiff-OR samples from the models stored in ``checkpoints``
    (exported parameter trees, see checkpoint.py), written as samples_{i}.npz.  FID itself (:253-278) needs the reference's
    TensorFlow Inception graph and dataset statistics and stays with the reference's evaluation.py."""
    params_list = [checkpoint.validate_params(checkpoint.load_params(c), config) for c in checkpoints]
    return evaluate_joint_samples(config, workdir, eval_folder, params_list, stoch=stoch, num_batches=num_batches, dt=dt,
                                  device=device, mode="or")


def evaluate_fid(config, workdir, eval_folder, stoch, checkpoint_path=None, num_batches=None, dt=None, device=None):
    """cifar/run_lib.py:129-198 up to the Inception stage: single-model samples through get_avg_vf([model], stoch).
    The reference restores the latest checkpoint of ``workdir`` (:133); here ``checkpoint_path`` (default
    ``<workdir>/params_ema.npz``) names the exported parameter tree."""
    path = checkpoint_path or os.path.join(workdir, "params_ema.npz")
    params = checkpoint.validate_params(checkpoint.load_params(path), config)
    return evaluate_joint_samples(config, workdir, eval_folder, [params], stoch=stoch, num_batches=num_batches, dt=dt,
                                  device=device, mode="avg")
