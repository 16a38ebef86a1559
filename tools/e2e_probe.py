import sys, time, torch
sys.path.insert(0, '/root/repo')
from super_diffusion_b200 import dynamics, eval_utils
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import utils as mutils
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
cfg = vpsde.get_config(); B=512; K=30
models, states = [], []
for m in range(2):
    model, params = mutils.init_model(10+m, cfg, zero_init_scale=1.0)
    models.append(model); states.append(mutils.State(params_ema=params, model_params=params))
cfg.eval.batch_size = B
vf = dynamics.get_joint_stoch_vf(0, models, states)
gen = eval_utils.get_generator(models, cfg, vf, dt=1.0/K - 1e-12, device=dev, return_logq=True)
shape=(B,32,32,3)
noise_host=[torch.randn(shape).pin_memory() for _ in range(4)]
trace=torch.empty(K,B,2).pin_memory(); x_host=torch.empty(shape).pin_memory()
nk=lambda i: noise_host[i%4]
gen(1,None,noise=nk,logq_trace=trace); torch.cuda.synchronize()
def ev(): e=torch.cuda.Event(enable_timing=True); e.record(); return e
for rep in range(3):
    t0=time.perf_counter(); s=ev()
    xg,ng,lq=gen(2,None,noise=nk,logq_trace=trace)
    t1=time.perf_counter()
    x_host.copy_(xg,non_blocking=True); e=ev(); torch.cuda.synchronize(); t2=time.perf_counter()
    print('call', ng, 'gpu ms total', s.elapsed_time(e), 'per step', s.elapsed_time(e)/ng, 'host submit ms', (t1-t0)*1e3, 'wall ms', (t2-t0)*1e3)
# components
torch.cuda.synchronize(); s=ev()
g=torch.Generator(device=dev); g.manual_seed(5); x=torch.randn(shape,generator=g,device=dev); e=ev(); torch.cuda.synchronize(); print('x0 draw gpu ms', s.elapsed_time(e))
t0=time.perf_counter(); g=torch.Generator(device=dev); g.manual_seed(5); x=torch.randn(shape,generator=g,device=dev); torch.cuda.synchronize(); print('x0 draw wall ms', (time.perf_counter()-t0)*1e3)
