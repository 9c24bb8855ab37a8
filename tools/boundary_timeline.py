"""Phase timeline of nfdpm_flow_boundary per level (nfdpm_flow_boundary_debug), BASELINE config-2 shapes, B = 128."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
from normalizing_flow import _native as N

dev = torch.device("cuda")
B = 128
for lvl, (C, hw) in enumerate([(12, 16), (24, 8), (48, 4)]):
    P = hw * hw
    M = B * P
    K1p = (9 * (C // 2) + 63) // 64 * 64
    ldp = (9 * C + 15) // 16 * 16
    pm = torch.randn(M, ldp, device=dev) * 0.3
    x = torch.randn(B, C, hw, hw, device=dev)
    y = torch.empty_like(x)
    a1 = torch.empty(M, K1p, dtype=torch.bfloat16, device=dev)
    mt, beta = torch.randn(C, C, device=dev) * 0.3, torch.randn(C, device=dev)
    b3, l3 = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    part = torch.empty(B, device=dev)
    run = lambda: N.flow_boundary(x, C * P, False, pm, ldp, b3, l3, part, mt, beta, y, C * P, a1, K1p, B, C, hw, hw, False)
    for _ in range(3):
        run()
    buf = torch.zeros(B, 16, dtype=torch.int64, device=dev)
    N.lib.nfdpm_flow_boundary_debug(buf.data_ptr())
    run(); run()
    torch.cuda.synchronize()
    N.lib.nfdpm_flow_boundary_debug(None)
    t = buf.cpu().double()
    d = (t[:, 1:8] - t[:, 0:7]) / 1.9e3
    names = ["stage params+x", "coupling", "log-det reduce(w0)", "(xs stash)", "mix", "y store", "im2col"]
    print(f"level {lvl} C={C} P={P}: total {float((t[:, 7] - t[:, 0]).mean()) / 1.9e3:.2f} us |",
          " | ".join(f"{n} {float(d[:, i].mean()):.2f}" for i, n in enumerate(names)))
