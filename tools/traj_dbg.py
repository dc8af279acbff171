import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import super_diff_disease_b200 as S
from oracle import superdiff_oracle as O
dev = torch.device("cuda:0")
T, shape = 4, (2, 1, 256, 256)
params = [O.init_unet_params(0), O.init_unet_params(1)]
models = []
for p in params:
    m = S.UNet(); m.load_state_dict(p); models.append(m.to(dev))
g = torch.Generator().manual_seed(shape[-1] + T)
stack = torch.randn((T,) + shape, generator=g)
xr, kr, lr = O.superposed_sample(params, O.Schedule(T), stack)
x, kap, lq = S.superposed_sample(models, S.DDPM(T), shape, dev, noise=stack.to(dev), return_trajectory=True)
torch.set_printoptions(precision=4, sci_mode=False, linewidth=200)
print("oracle logq\n", lr); print("gpu logq\n", lq.cpu()); print("oracle kappa\n", kr); print("gpu kappa\n", kap.cpu())
