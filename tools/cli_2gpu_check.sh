set -e
python - <<'PY'
import torch
from super_diffusion_b200 import checkpoint
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import ddpm, utils as mutils
cfg = vpsde.get_config()
for s, n in ((1, "a.npz"), (2, "b.npz")):
    checkpoint.save_npz("/tmp/" + n, mutils.init_model(s, cfg, zero_init_scale=1.0)[1])
print("saved")
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 -m super_diffusion_b200.main --config vpsde --workdir /tmp/w --mode eval_joint_fid_stoch --chkpts /tmp/a.npz,/tmp/b.npz --batch_size 16 --num_batches 2 --dt 0.25
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 -m super_diffusion_b200.main --config vpsde --workdir /tmp/w --mode eval_joint_fid --chkpts /tmp/a.npz,/tmp/b.npz --batch_size 16 --num_batches 1 --dt 0.25
python - <<'PY'
import numpy as np, os
for d in ("samples_stoch", "samples"):
    p = f"/tmp/w/eval/{d}"
    fs = sorted(os.listdir(p)); z = np.load(os.path.join(p, fs[0]))
    print(d, fs, z["samples"].shape, z["samples"].dtype, int(z["num_steps"]), z["samples"].mean())
PY
