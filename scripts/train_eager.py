"""A few eager (not graph-replayed) training iterations of BASELINE config #3, for `ncu -k regex:...` captures:
    python scripts/train_eager.py [iters]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import danbo_b200 as db
from danbo_b200 import synthetic as syn, skeleton as sk, training
dev = torch.device("cuda", 0)
args = db.make_args("danbo_cfg3", no_reload=True)
data_attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
_, kw_test, *_ = db.create_raycaster(args, data_attrs, device=dev)
caster = kw_test["ray_caster"]; caster.network.load_state_dict(syn.synthetic_params(0))
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.training_batch(16, 192, seed=0).items()}
step = training.TrainStep(caster, args)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    loss, _ = step(batch)
torch.cuda.synchronize()
print("loss", float(loss))
