"""One eager forward (+ inverse) of a BASELINE config between cudaProfilerStart/Stop, for ncu launch lists:
    NFDPM_GRAPHS=0 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_cfg4.csv python tools/run_cfg.py 4 bf16"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
cfg = int(sys.argv[1]); mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
os.environ["NFDPM_PRECISION"] = mode
os.environ.setdefault("NFDPM_GRAPHS", "0")
import torch
import bench_configs as BC
import normalizing_flow as nf
c, L, K, B, S, _ = BC.CONFIGS[cfg]
flow, prior, x = BC.build(c, L, K, B, S)
with torch.no_grad():
    def step():
        ld, lp = nf.initialize_with_zeros(2, B, BC.DEV)
        zs, ld, lp = flow.transform(x, ld, lp)
        lp += prior.compute_log_prob(zs[-1])
        return flow.invert(zs)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok")
