import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_reference_vectors as T
from super_diffusion_b200.models import utils as mutils
cuda = torch.device("cuda:0")
for name, c in T._load("ref_scorenet.npz").items():
    config, model, params = T._our_params(c)
    fn = mutils.get_model_fn(model, params)
    out = fn(T._t(c["t"]).to(cuda), T._t(c["x"]).to(cuda).contiguous(), T._t(c["y"]).to(cuda))
    got, ref = out.double().cpu().numpy(), c["out"]
    print(name, "rel-RMS", float(np.sqrt(((got - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean())), "max err / max", float(np.abs(got - ref).max() / np.abs(ref).max()))
