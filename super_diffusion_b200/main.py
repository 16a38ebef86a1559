"""Command line — mirror of the reference's cifar/main.py (:10-40) for the sampling modes.

    python -m super_diffusion_b200.main --config vpsde --workdir W --mode eval_joint_fid_stoch --chkpts a.npz,b.npz
    torchrun --nproc-per-node 8 -m super_diffusion_b200.main ...        # batch sharded over the GPUs of one box

``--config`` is a name from super_diffusion_b200/configs (vpsde, vpsdeA, vpsdeB) or a path to a Python file with
``get_config()`` (the reference's cifar/configs/sm/cifar/*.py load unchanged if ml_collections is importable).
Modes ``train`` and ``fid_stats`` belong to training / FID (outside the sampling path) and are rejected.
"""
import argparse
import importlib.util
import os

import torch

from . import run_lib
from .configs import vpsde

MODES = ["train", "eval_fid", "eval_fid_stoch", "eval_joint_fid", "eval_joint_fid_stoch", "fid_stats"]


def load_config(name):
    if os.path.isfile(name):
        spec = importlib.util.spec_from_file_location("sd_user_config", name)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.get_config()
    table = {"vpsde": dict(), "vpsdeA": dict(conditioned=True, train_split="train[:50%]"),
             "vpsdeB": dict(conditioned=True, train_split="train[50%:]")}
    if name not in table:
        raise ValueError(f"unknown config {name!r}: expected one of {sorted(table)} or a path to a config file")
    return vpsde.get_config(**table[name])


def build_parser():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", required=True)
    ap.add_argument("--workdir", required=True)
    ap.add_argument("--mode", required=True, choices=MODES)
    ap.add_argument("--eval_folder", default="eval")
    ap.add_argument("--chkpts", default=None, help="paths to exported parameter trees for joint evaluation (comma separated)")
    ap.add_argument("--num_batches", type=int, default=None, help="stop after this many batches (default: eval.num_samples)")
    ap.add_argument("--batch_size", type=int, default=None, help="override config.eval.batch_size")
    ap.add_argument("--dt", type=float, default=None, help="override the reference's dt = 5e-3 (cifar/eval_utils.py:75)")
    ap.add_argument("--precision", default=None, choices=["bf16", "fp32"],
                    help="score-net arm: bf16 (bf16 operands / activations, fp32 accumulation; default) or fp32 (FP32-faithful: the "
                         "reference's fp32 arithmetic to ~1e-5, ~2.6x slower); sets config.model.precision")
    return ap


def launch(argv=None):
    args = build_parser().parse_args(argv)
    if args.mode in ("train", "fid_stats"):
        raise SystemExit(f"mode {args.mode!r} is outside the sampling path this package accelerates (use the reference)")
    config = load_config(args.config)
    if args.batch_size:
        config.eval.batch_size = args.batch_size
    if args.precision:
        config.model.precision = args.precision
    if "LOCAL_RANK" in os.environ and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        torch.distributed.init_process_group("nccl")
    stoch = args.mode.endswith("_stoch")
    if args.mode.startswith("eval_joint_fid"):
        if not args.chkpts:
            raise SystemExit("--chkpts is required for joint evaluation")
        chk = [c.strip() for c in args.chkpts.split(",")]
        out = run_lib.evaluate_joint_fid(config, args.workdir, args.eval_folder, chk, stoch, num_batches=args.num_batches, dt=args.dt)
    else:
        out = run_lib.evaluate_fid(config, args.workdir, args.eval_folder, stoch, num_batches=args.num_batches, dt=args.dt)
    if torch.distributed.is_initialized():
        torch.distributed.barrier()
        if torch.distributed.get_rank() == 0:
            print(out)
        torch.distributed.destroy_process_group()
    else:
        print(out)
    return out


if __name__ == "__main__":
    launch()
