"""sd_scorenet_forward: the whole score network (cifar/models/ddpm.py:47-101, the reference's model_fn seam of
cifar/models/utils.py:86-96) as ONE native call.  CPU: the C++ layout walk and the Python packer agree on the weight blob, the
workspace dry run, argument validation.  GPU: bit-identical to the Python-driven forward (same kernels, same order), and within
the bf16 bound of the fp64 oracle."""
import ctypes

import pytest
import torch

from super_diffusion_b200 import _lib, native
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import utils as mutils


def _cfgs():
    return [("vpsde", vpsde.get_config()), ("vpsdeA", vpsde.get_config(conditioned=True))]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("name,cfg", _cfgs())
def test_blob_layout_matches_the_native_walk(lib, name, cfg, precision):
    model, params = mutils.init_model(3, cfg, zero_init_scale=1.0)
    bound = model.bind(params, torch.device("cpu"), precision=precision)          # weight preparation is plain tensor code
    blob = native.pack_weights(bound)
    desc = native.make_desc(cfg, precision=native.PRECISION_FP32_FAITHFUL if precision == "fp32" else native.PRECISION_BF16)
    n = ctypes.c_size_t()
    assert lib.sd_scorenet_weights_bytes(ctypes.byref(desc), ctypes.byref(n)) == 0
    assert n.value == blob.numel() and n.value % native.ALIGN == 0
    # 36.0 M parameters: bf16 GEMM weights (hi|lo pairs in the FP32-faithful arm) + fp32 vectors + the
    # identity / padding segments
    k = 2 if precision == "fp32" else 1
    assert k * 70e6 < n.value < k * 80e6


def test_workspace_dry_run_and_validation(lib):
    cfg = vpsde.get_config()
    desc = native.make_desc(cfg)
    sizes = []
    for B in (8, 64, 512):
        w = ctypes.c_size_t()
        assert lib.sd_scorenet_workspace_bytes(ctypes.byref(desc), B, 0, ctypes.byref(w)) == 0
        sizes.append(w.value)
    assert sizes[0] < sizes[1] < sizes[2] < 16 * 2 ** 30
    assert 6.5 < sizes[2] / sizes[1] < 8.5                      # activations scale with the batch
    w = ctypes.c_size_t()
    # the 4x4 mid-block attention packs 8 images per 128-row tile: smaller / ragged batches are padded internally, so every
    # activation has room for the batch rounded up to 8 images
    assert lib.sd_scorenet_workspace_bytes(ctypes.byref(desc), 2, 0, ctypes.byref(w)) == 0
    assert 0 < w.value <= sizes[0]
    bad = native.make_desc(cfg)
    bad.nf = 100
    assert lib.sd_scorenet_weights_bytes(ctypes.byref(bad), ctypes.byref(w)) == -2
    null = ctypes.c_void_p(0)
    assert lib.sd_scorenet_forward(ctypes.byref(desc), null, 0, null, null, 8, null, null, 0, 1, null) == -1
    assert lib.sd_scorenet_forward(ctypes.byref(desc), null, 0, null, null, 8, null, null, 0, 0, null) == -2      # precision
    assert lib.sd_scorenet_forward(ctypes.byref(desc), null, 0, null, null, 8, null, null, 0, 2, null) == -1      # blob is bf16
    assert b"precision" in lib.sd_last_error()
    assert lib.sd_scorenet_forward(ctypes.byref(desc), null, 0, null, null, 0, null, null, 0, 1, null) == 0       # empty batch


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("name,cfg", _cfgs())
def test_native_forward_is_the_python_forward(cuda, name, cfg, precision):
    model, params = mutils.init_model(7, cfg, zero_init_scale=1.0)
    bound = model.bind(params, cuda, precision=precision)
    net = native.NativeScoreNet(bound)
    B = 16
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 32, 32, 3, generator=g).to(cuda)
    y = (torch.arange(B) % 10).to(cuda) if cfg.model.conditioned else None
    for t in (0.6, torch.linspace(0.05, 0.95, B)):
        want = bound(t if not torch.is_tensor(t) else t.to(cuda), x, y)
        got = net(t, x, y)
        torch.cuda.synchronize()
        assert torch.equal(got, want), (got - want).abs().max()
    # a second call reuses the workspace and gives the same bits
    assert torch.equal(net(0.6, x, y), bound(0.6, x, y))


@pytest.mark.gpu
def test_native_forward_against_the_oracle(cuda):
    from oracle import scorenet as OS
    cfg = vpsde.get_config()
    model, params = mutils.init_model(5, cfg, zero_init_scale=1.0)
    net = native.NativeScoreNet(model.bind(params, cuda))
    B, t = 8, 0.4
    x = torch.randn(B, 32, 32, 3, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = OS.scorenet_apply(params, cfg, torch.full((B, 1, 1, 1), t), x, None)
    got = net(t, x.to(cuda)).cpu()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 1.9e-2, rel          # bf16 operands / activations, fp32 accumulation (stated separately from the fp32 step bound)


@pytest.mark.gpu
def test_native_forward_errors(cuda):
    cfg = vpsde.get_config()
    model, params = mutils.init_model(5, cfg, zero_init_scale=1.0)
    net = native.NativeScoreNet(model.bind(params, cuda))
    x = torch.randn(8, 32, 32, 3, device=cuda)
    # a batch that does not fill the packed low-resolution attention tiles (4x4: 8 images per tile) is padded internally
    # (the reference's eval batch is 100): same scores as inside the full batch
    full = net(0.5, x)
    part = net(0.5, x[:2].contiguous())
    assert torch.allclose(full[:2], part, rtol=0, atol=1e-6 + 1e-3 * full.abs().max().item())
    lib = _lib.load()
    out = torch.empty_like(x)
    t = torch.full((1,), 0.5, device=cuda)
    small = torch.empty(1 << 20, dtype=torch.uint8, device=cuda)
    rc = lib.sd_scorenet_forward(ctypes.byref(net.desc), t.data_ptr(), 0, x.data_ptr(), None, 8, out.data_ptr(), small.data_ptr(),
                                 small.numel(), 1, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == -1 and b"workspace" in lib.sd_last_error()
    d2 = native.make_desc(cfg, net.blob)
    d2.weights_bytes = net.blob.numel() - 256
    rc = lib.sd_scorenet_forward(ctypes.byref(d2), t.data_ptr(), 0, x.data_ptr(), None, 8, out.data_ptr(), small.data_ptr(),
                                 small.numel(), 1, torch.cuda.current_stream().cuda_stream)
    assert rc == -1 and b"weights_bytes" in lib.sd_last_error()


@pytest.mark.gpu
def test_plain_c_host_gets_the_same_scores(cuda, tmp_path):
    """examples/native_forward.c: gcc + libsuperdiff_b200.so + the CUDA runtime, no Python in the process - fed the saved weight
    blob and a raw input file, it must write the bits the Python-driven forward produces."""
    import os
    import shutil
    import subprocess
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this box")
    exe = str(tmp_path / "native_forward")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(root, "examples", "native_forward.c"), "-o", exe,
                    "-L", os.path.join(root, "super_diffusion_b200"), "-lsuperdiff_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
                    "-Wl,-rpath," + os.path.join(root, "super_diffusion_b200")], check=True)
    cfg = vpsde.get_config()
    model, params = mutils.init_model(9, cfg, zero_init_scale=1.0)
    bound = model.bind(params, cuda)
    net = native.NativeScoreNet(bound)
    net.save(str(tmp_path / "model"))
    B, t = 8, 0.35
    x = torch.randn(B, 32, 32, 3, generator=torch.Generator().manual_seed(4))
    x.numpy().tofile(str(tmp_path / "x.bin"))
    r = subprocess.run([exe, str(tmp_path / "model.bin"), str(tmp_path / "x.bin"), str(tmp_path / "out.bin"), str(B), str(t)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = torch.from_numpy(np.fromfile(str(tmp_path / "out.bin"), dtype=np.float32).reshape(B, 32, 32, 3))
    want = bound(t, x.to(cuda)).cpu()
    assert torch.equal(got, want), (got - want).abs().max()


@pytest.mark.gpu
def test_plain_c_sampler_matches_the_python_loop(cuda, tmp_path):
    """examples/native_sampler.c: the SuperDiff-OR loop (cifar/eval_utils.py:72-86 over cifar/dynamics.py:115-136) from C - per
    timestep two sd_scorenet_forward calls and one sd_step_vpsde.  Same noise in, same bits out as the Python-driven loop."""
    import os
    import shutil
    import subprocess
    import numpy as np
    from super_diffusion_b200 import ops, sde
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this box")
    exe = str(tmp_path / "native_sampler")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(root, "examples", "native_sampler.c"), "-o", exe,
                    "-L", os.path.join(root, "super_diffusion_b200"), "-lsuperdiff_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
                    "-Wl,-rpath," + os.path.join(root, "super_diffusion_b200")], check=True)
    cfg = vpsde.get_config()
    B, n_steps, dt, M = 8, 4, 1e-3, 2
    bounds = []
    for m in range(M):
        model, params = mutils.init_model(20 + m, cfg, zero_init_scale=1.0)
        b = model.bind(params, cuda)
        bounds.append(b)
        native.NativeScoreNet(b).save(str(tmp_path / f"model{m}"))
    g = torch.Generator().manual_seed(8)
    x0 = torch.randn(B, 32, 32, 3, generator=g)
    noise = torch.randn(n_steps, B, 32, 32, 3, generator=g)
    x0.numpy().tofile(str(tmp_path / "x0.bin"))
    noise.numpy().tofile(str(tmp_path / "noise.bin"))
    r = subprocess.run([exe, str(tmp_path / "x0.bin"), str(tmp_path / "noise.bin"), str(tmp_path / "x.bin"), str(tmp_path / "logq.bin"),
                        str(B), str(n_steps), repr(dt)] + [str(tmp_path / f"model{m}.bin") for m in range(M)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got_x = torch.from_numpy(np.fromfile(str(tmp_path / "x.bin"), dtype=np.float32).reshape(B, 32, 32, 3))
    got_l = torch.from_numpy(np.fromfile(str(tmp_path / "logq.bin"), dtype=np.float32).reshape(B, M))
    # the same calls from Python
    x = x0.to(cuda).clone()
    logq = torch.zeros(B, M, device=cuda)
    t = 1.0
    for i in range(n_steps):
        tt = torch.full((1,), t, device=cuda, dtype=torch.float32)
        scores = [b(tt, x) for b in bounds]
        ops.step_vpsde(x, noise[i].to(cuda), scores, logq, sde.dlog_alphadt(t), sde.beta(t), sde.sigma(t), dt, ops.MODE_OR,
                       ops.DLOGQ_CIFAR_MAXSUB, temperature=1e6, x_out=x)
        t += -dt
    torch.cuda.synchronize()
    assert torch.equal(got_x, x.cpu()), (got_x - x.cpu()).abs().max()
    assert torch.equal(got_l, logq.cpu())


@pytest.mark.gpu
def test_native_sampling_loop_as_one_cuda_graph(cuda):
    """sd_scorenet_forward_sched + sd_step_vpsde(sched, counter) + sd_counter_add captured ONCE and replayed for every timestep:
    the same samples and log-densities as the eager loop with host-side scalars."""
    from super_diffusion_b200 import ops, sde
    cfg = vpsde.get_config()
    B, n_steps, dt, M = 8, 5, 1e-3, 2
    nets, bounds = [], []
    for m in range(M):
        model, params = mutils.init_model(30 + m, cfg, zero_init_scale=1.0)
        b = model.bind(params, cuda)
        bounds.append(b)
        nets.append(native.NativeScoreNet(b))
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(B, 32, 32, 3, generator=g).to(cuda)
    noise = torch.randn(n_steps, B, 32, 32, 3, generator=g).to(cuda)
    ts = sde.time_grid(n_steps, dt)
    sched = sde.schedule_table(ts, dt, cuda)
    counter = torch.zeros(1, dtype=torch.int32, device=cuda)
    x, nz = x0.clone(), torch.empty_like(x0)
    logq, w = torch.zeros(B, M, device=cuda), torch.zeros(B, M, device=cuda)
    scores = [torch.empty_like(x0) for _ in range(M)]

    def step():
        for m in range(M):
            nets[m].forward_sched(sched, counter, x, out=scores[m])
        ops.step_vpsde(x, nz, scores, logq, 0.0, 0.0, 1.0, 0.0, ops.MODE_OR, ops.DLOGQ_CIFAR_MAXSUB, temperature=1e6, x_out=x,
                       weights=w, sched=sched, step_counter=counter)
        ops.counter_add(counter, 1)
    nz.copy_(noise[0])
    side = torch.cuda.Stream(device=cuda)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()                                  # warm-up outside the capture (workspaces, function attributes)
    torch.cuda.current_stream().wait_stream(side)
    x.copy_(x0); logq.zero_(); counter.zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    x.copy_(x0); logq.zero_(); counter.zero_()
    for i in range(n_steps):
        nz.copy_(noise[i])
        graph.replay()
    torch.cuda.synchronize()
    # eager loop, scalars by value, Python-driven forward
    xr, lr = x0.clone(), torch.zeros(B, M, device=cuda)
    for i in range(n_steps):
        t = float(ts[i])
        sc = [b(torch.full((1,), t, device=cuda, dtype=torch.float32), xr) for b in bounds]
        ops.step_vpsde(xr, noise[i], sc, lr, sde.dlog_alphadt(t), sde.beta(t), sde.sigma(t), dt, ops.MODE_OR,
                       ops.DLOGQ_CIFAR_MAXSUB, temperature=1e6, x_out=xr)
    torch.cuda.synchronize()
    assert int(counter.item()) == n_steps
    assert torch.allclose(x, xr, rtol=1e-5, atol=1e-5), (x - xr).abs().max()
    assert torch.allclose(logq, lr, rtol=1e-4, atol=1e-3), (logq - lr).abs().max()
