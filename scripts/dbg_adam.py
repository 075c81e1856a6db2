import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
import danbo_b200 as db
from danbo_b200 import synthetic as syn, skeleton as sk, training
dev = torch.device("cuda", 0)
def run(kind, graph):
    args = db.make_args("danbo_cfg3", no_reload=True)
    attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
    _, kw, *_ = db.create_raycaster(args, attrs, device=dev)
    caster = kw["ray_caster"]; caster.network.load_state_dict(syn.synthetic_params(0))
    full = syn.training_batch(16, 192, seed=0)
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in full.items()}; batch["N_uniques"] = 16
    params = [p for p in caster.network.parameters() if p.requires_grad]
    opt = None if kind == "flat" else torch.optim.Adam(params, lr=args.lrate, betas=(0.9, 0.999), fused=True, capturable=graph)
    step = training.TrainStep(caster, args, optimizer=opt, graph=graph)
    torch.manual_seed(7)
    out = []
    for i in range(12):
        loss, _ = step(batch); out.append(float(loss))
    return out
for graph in (False, True):
    a, b = run("torch", graph), run("flat", graph)
    print("graph", graph)
    print(" torch", " ".join(f"{x:.4f}" for x in a))
    print(" flat ", " ".join(f"{x:.4f}" for x in b))
