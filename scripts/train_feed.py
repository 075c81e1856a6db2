"""Training iterations/s with the data feed in the loop (and optionally the pose layer), on synthetic data:

    python scripts/train_feed.py [--iters 200] [--images 64] [--size 512] [--opt_pose] [--graph]
    torchrun --nproc-per-node 2 scripts/train_feed.py ...      (images of a batch dealt to the ranks)

Config #3's step (danbo_base, 64 + 16 samples, 16 images x 192 rays, Adam) fed by `feed.RayFeed` instead of a fixed
batch: what a trainer that replaces the reference's h5py DataLoader would see.  NOT YET RUN ON HARDWARE (written after
round 1's GPU minutes were spent); not a bench value."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                          # noqa: E402
import danbo_b200 as db                               # noqa: E402
from danbo_b200 import feed as fd, parallel, pose_opt as po, skeleton as sk, synthetic as syn, training  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--opt_pose", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the iteration from a CUDA graph (not with --opt_pose)")
    ap.add_argument("--prefetch", action="store_true", help="draw the next batch on a side stream during the iteration")
    a = ap.parse_args()
    rank, world, local = parallel.init_distributed() if "RANK" in os.environ else (0, 1, 0)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    args = db.make_args("danbo_cfg3", no_reload=True)
    attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
    _, kw_test, *_ = db.create_raycaster(args, attrs, device=dev)
    caster = kw_test["ray_caster"]
    caster.network.load_state_dict(syn.synthetic_params(0))
    arrays = fd.synthetic_arrays(n_images=a.images, H=a.size, W=a.size, seed=0)
    feed = fd.RayFeed.from_arrays(arrays, syn.NEAR, syn.FAR, N_rand=3072, N_sample_images=16, device=dev, rank=rank,
                                  world_size=world, cam_idxs=torch.arange(a.images) % 8)
    popt_kw = pose_optimizer = None
    if a.opt_pose:
        args.opt_pose_coef, args.opt_pose_tol = 2.0, 0.0
        pose_optimizer, popt_kw = po.create_popt(args, {"rest_pose": syn.rest_pose()[None], "betas": torch.zeros(1, 10).numpy(),
                                                        "kp3d": arrays["kp3d"], "bones": arrays["bones"]}, device=dev)
    step = training.TrainStep(caster, args, world_size=world, graph=a.graph and not a.opt_pose, popt_kwargs=popt_kw,
                              pose_optimizer=pose_optimizer)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    draw = feed.prefetch if a.prefetch else feed.next_batch
    for _ in range(5):
        loss, _ = step(draw(gen))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        loss, _ = step(draw(gen))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    if rank == 0:
        print(f"{1e3 / ms:.1f} it/s  {ms:.3f} ms/iter (device)  {(time.perf_counter() - t0) * 1e3 / a.iters:.3f} ms/iter (wall)  "
              f"loss {float(loss):.4f}  world {world}  opt_pose {a.opt_pose}  graph {a.graph and not a.opt_pose}  prefetch {a.prefetch}")


if __name__ == "__main__":
    main()
