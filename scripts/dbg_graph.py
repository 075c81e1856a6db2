import sys, os, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
from util import make_caster
from danbo_b200 import synthetic as syn, training
DEV = "cuda"
caster, args, _ = make_caster("danbo_cfg3", train=True)
batch = syn.training_batch(4, 48, seed=2)
batch = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()}
step = training.TrainStep(caster, args, graph=False)
for _ in range(3): step(batch)
torch.cuda.synchronize()
def fwd():
    step.bucket.zero()
    preds = caster(batch["ray_batch"], N_samples=args.N_samples, kp_batch=batch["kp_batch"], skts=batch["skts"], cyls=batch["cyls"], bones=batch["bones"], cams=batch["cams"], N_uniques=batch["N_uniques"], perturb=1.0, N_importance=args.N_importance, raw_noise_std=1.0)
    loss, _ = training.compute_loss(args, preds, batch, caster.network)
    return loss
for name, fn in (("fwd", lambda: fwd()), ("fwd+bwd", lambda: fwd().backward()), ("full", lambda: step._step(batch))):
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "capture OK")
    except Exception as e:
        print(name, "FAILED"); traceback.print_exc(limit=12)
        break
