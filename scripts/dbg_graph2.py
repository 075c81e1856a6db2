import sys, os, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
from util import make_caster
from danbo_b200 import synthetic as syn, training, kernels as K, autograd as ag
DEV = "cuda"
caster, args, _ = make_caster("danbo_cfg3", train=True)
batch = syn.training_batch(4, 48, seed=2)
batch = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()}
step = training.TrainStep(caster, args, graph=False)
for _ in range(2): step(batch)
torch.cuda.synchronize()
# monkeypatch every kernel wrapper used in backward to report capture status after the call
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
def status(tag):
    st = ctypes.c_int(0)
    rc = rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    print(f"   after {tag}: rc={rc} capture_status={st.value}", flush=True)
for name in ("merge_composite_bwd", "composite_bwd", "mlp_backward", "field_agg_bwd", "ray_bias_bwd", "_dgrad", "_wgrad", "_colsum"):
    orig = getattr(K, name)
    def wrap(*a, _o=orig, _n=name, **k):
        r = _o(*a, **k); status(_n); return r
    setattr(K, name, wrap)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        step.bucket.zero()
        preds = caster(batch["ray_batch"], N_samples=args.N_samples, kp_batch=batch["kp_batch"], skts=batch["skts"], cyls=batch["cyls"], bones=batch["bones"], cams=batch["cams"], N_uniques=batch["N_uniques"], perturb=1.0, N_importance=args.N_importance, raw_noise_std=1.0)
        loss, _ = training.compute_loss(args, preds, batch, caster.network)
        status("forward")
        loss.backward()
except Exception as e:
    print("FAILED:", str(e)[:200])
