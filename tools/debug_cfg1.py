"""Localise the cfg1 inversion NaN: compare every image / every stage against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
import normalizing_flow as nf
from oracle import glow_oracle as O
torch.set_grad_enabled(False)
DEV = torch.device("cuda")
c, L, K, B, S = 1, 3, 4, 64, 32
torch.manual_seed(0)
flow = nf.Glow(c, L, K).to(DEV)
x = O.seeded_input((B, c, S, S), 123).to(DEV)
ld = torch.zeros(B, dtype=torch.float64, device=DEV); lp = torch.zeros(B, dtype=torch.float64, device=DEV)
zs, ld, lp = flow.transform(x, ld, lp)
sd = {k: v.cpu() for k, v in flow.state_dict().items()}
g = torch.Generator().manual_seed(1)
for k in list(sd):
    if k.endswith("net.4.weight") or k.endswith("net.4.bias") or k.endswith("net.4.logs") or ".split.conv." in k:
        sd[k] = sd[k] + 0.01 * torch.randn(sd[k].shape, generator=g)
flow.load_state_dict(sd)
ld = torch.zeros(B, dtype=torch.float64, device=DEV); lp = torch.zeros(B, dtype=torch.float64, device=DEV)
zs, ld, lp = flow.transform(x, ld, lp)
ld_o = torch.zeros(B, dtype=torch.float64); lp_o = torch.zeros(B, dtype=torch.float64)
zo, ld_o, lp_o = O.glow_transform(sd, x.cpu(), L, K, ld_o, lp_o)
for i, (a, b) in enumerate(zip(zs, zo)):
    d = (a.cpu() - b).flatten(1).abs().max(1)[0]
    print(f"fwd z{i}: finite={torch.isfinite(a).all().item()} worst img={int(d.argmax())} maxerr={d.max().item():.3e}")
print("ld err", (ld.cpu() - ld_o).abs().max().item(), "lp err", (lp.cpu() - lp_o).abs().max().item())
# inverse from the ORACLE latents, stage by stage
lat = [t.to(DEV) for t in zo]
xr = flow.invert(lat)
bad = (~torch.isfinite(xr)).flatten(1).any(1)
print("invert(all latents): nonfinite images:", bad.nonzero().flatten().tolist()[:20], "of", B)
xo = O.glow_invert(sd, zo, L, K)
good = ~bad.cpu()
if good.any():
    print("max err on finite images", (xr.cpu()[good] - xo[good]).abs().max().item())
# final flows only
y = lat[-1]
yo = zo[-1]
for j, st in enumerate(reversed(flow.final_flows)):
    y = st.invert(y)
    yo = O.step_inv(yo, sd, f"final_flows.{K - 1 - j}.")
    e = (y.cpu() - yo).flatten(1).abs().max(1)[0]
    print(f"final step {K-1-j}: finite={torch.isfinite(y).all().item()} worst={int(e.argmax())} err={e.max().item():.3e}")
# twice in a row (workspace reuse)
xr2 = flow.invert(lat)
print("second invert finite:", torch.isfinite(xr2).all().item(), "equal to first where finite:",
      torch.equal(torch.nan_to_num(xr), torch.nan_to_num(xr2)))
xr3 = flow.invert([t.clone() for t in zs])
print("invert(own zs) finite:", torch.isfinite(xr3).all().item(), (xr3 - x).abs().max().item())
