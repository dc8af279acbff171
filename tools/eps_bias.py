"""Is the UNet-forward error vs the fp32 oracle systematic (a gain error) or random?  y ~ s * ref + residual."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import super_diff_disease_b200 as S
from oracle import superdiff_oracle as O
dev = torch.device("cuda:0")
for wseed in (0, 1):
    p = O.init_unet_params(wseed)
    m = S.UNet(); m.load_state_dict(p); m = m.to(dev)
    for R, t in [(64, 3), (128, 2), (256, 3), (256, 0)]:
        g = torch.Generator().manual_seed(R + t)
        x = torch.randn((2, 1, R, R), generator=g)
        tt = torch.full((2,), t, dtype=torch.long)
        with torch.no_grad():
            ref = O.unet_forward(p, x, tt)
        y = m(x.to(dev), tt.to(dev)).cpu()
        s = (y * ref).sum() / (ref * ref).sum()
        res = y - s * ref
        print(f"w{wseed} R={R} t={t}: rel-L2 {((y-ref).norm()/ref.norm()).item():.4e}  gain s={s.item():.6f}  residual rel-L2 {(res.norm()/ref.norm()).item():.4e}  mean(y)-mean(ref) {(y.mean()-ref.mean()).item():.3e} ref rms {ref.pow(2).mean().sqrt().item():.3e}")
