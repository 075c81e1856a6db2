"""The training data feed on the device (run with -m gpu): `feed.RayFeed(device="cuda")` replays the
image / pixel indices of the reference's own batches (tests/golden/feed_*.npz, made by the unmodified `BaseH5Dataset` +
`ray_collate_fn`) and must reproduce them exactly as the CPU feed does (tests/test_feed.py); its own on-device draws are
checked by their properties; and one training iteration runs on a batch that never left the GPU."""
import os

import numpy as np
import pytest
import torch

from danbo_b200 import feed as fd, synthetic as syn

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def _feed_for(fx, **kw):
    n, H, W, seed, centers = (int(v) for v in fx["spec"])
    arrays = fd.synthetic_arrays(n, H, W, seed, centers=bool(centers))
    assert int(arrays["imgs"].astype(np.int64).sum()) == int(fx["imgs_checksum"][0]), "synthetic data set drifted"
    return fd.RayFeed.from_arrays(arrays, syn.NEAR, syn.FAR, N_rand=48, N_sample_images=4, perturb_bg=False, device=DEV, **kw)


@pytest.mark.parametrize("name", ["feed_plain", "feed_centers"])
def test_device_feed_reproduces_reference_batch(name):
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    feed = _feed_for(fx)
    b = feed.next_batch(image_idxs=fx["image_idxs"][::-1].copy(), pixel_idxs=fx["pixel_idxs"])
    assert all(v.is_cuda for v in b.values() if torch.is_tensor(v))
    rays = b["ray_batch"].cpu().numpy()
    np.testing.assert_array_equal(rays[:, 0:3], fx["rays_o"])
    np.testing.assert_allclose(rays[:, 3:6], fx["rays_d"], rtol=0, atol=3e-7)       # fp32 dot product, summation order
    d = fx["rays_d"]
    np.testing.assert_allclose(rays[:, 8:11], d / np.linalg.norm(d, axis=-1, keepdims=True), atol=4e-7)
    for key, ref in (("target_s", "target_s"), ("fgs", "fgs"), ("bgs", "bgs"), ("kp_batch", "kp3d"), ("bones", "bones"),
                     ("skts", "skts"), ("cyls", "cyls"), ("kp_idx", "kp_idx")):
        np.testing.assert_array_equal(b[key].cpu().numpy(), fx[ref], err_msg=key)   # gathers and uint8 -> float: bit-exact
    np.testing.assert_array_equal(b["cams"][:, 0].cpu().numpy(), fx["cam_idxs"])
    assert b["N_uniques"] == 4


def test_device_draws_and_a_training_iteration():
    """On-device draws: distinct sorted images, R distinct in-mask pixels per image; then one TrainStep on that batch."""
    import danbo_b200 as db
    from danbo_b200 import skeleton as sk, training
    feed = fd.synthetic_feed(n_images=8, H=64, W=64, N_rand=4 * 48, N_sample_images=4, device=DEV)
    torch.manual_seed(0)
    b = feed.next_batch()
    img, pix = feed.last_idxs
    img, pix = img.cpu().numpy(), pix.cpu().numpy()
    assert len(set(img.tolist())) == 4 and (np.diff(img) > 0).all()
    assert pix.shape == (4, 48)
    masks = feed.sampling_masks.reshape(feed.sampling_masks.shape[0], -1).cpu().numpy()
    for k in range(4):
        assert len(set(pix[k].tolist())) == 48 and (np.diff(pix[k]) > 0).all()
        assert masks[img[k], pix[k]].all()
    args = db.make_args("danbo_fast", no_reload=True)
    attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
    _, kw, *_ = db.create_raycaster(args, attrs, device=DEV)
    caster = kw["ray_caster"]
    caster.network.load_state_dict(syn.synthetic_params(0))
    step = training.TrainStep(caster, args)
    losses = []
    for _ in range(6):
        loss, _ = step(feed.next_batch())
        losses.append(float(loss))
    print("[feed] losses with the device feed in the loop", ["%.4f" % l for l in losses])
    assert all(l == l for l in losses)
