"""Training data feed (`feed.RayFeed`) against the reference's data pipeline (core/dataset.py).

CPU only: the feed is torch ops on whatever device holds the arrays.  `tests/golden/feed_*.npz` are batches of the
UNMODIFIED reference `BaseH5Dataset` + `ray_collate_fn` on the synthetic training set (oracle/gen_golden_feed.py); the
feed replays the image / pixel indices the reference drew.  The random draws themselves are checked by their
properties."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import danbo_b200                                    # noqa: E402,F401
from danbo_b200 import feed as fd                    # noqa: E402
from danbo_b200 import synthetic as syn              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _feed_for(fx, **kw):
    n, H, W, seed, centers = (int(v) for v in fx["spec"])
    arrays = fd.synthetic_arrays(n, H, W, seed, centers=bool(centers))
    assert int(arrays["imgs"].astype(np.int64).sum()) == int(fx["imgs_checksum"][0]), "synthetic data set drifted"
    return fd.RayFeed.from_arrays(arrays, syn.NEAR, syn.FAR, N_rand=48, N_sample_images=4, perturb_bg=False, **kw), arrays


@pytest.mark.parametrize("name", ["feed_plain", "feed_centers"])
def test_feed_reproduces_reference_batch(name):
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    feed, _ = _feed_for(fx)
    b = feed.next_batch(image_idxs=fx["image_idxs"][::-1].copy(), pixel_idxs=fx["pixel_idxs"])   # unsorted: the feed sorts
    n = fx["rays_o"].shape[0]
    rays = b["ray_batch"].numpy()
    assert rays.shape == (n, 11)
    np.testing.assert_array_equal(rays[:, 0:3], fx["rays_o"])
    np.testing.assert_allclose(rays[:, 3:6], fx["rays_d"], rtol=0, atol=2e-7)       # fp32 dot product, summation order
    np.testing.assert_array_equal(rays[:, 6], np.full(n, syn.NEAR, np.float32))
    np.testing.assert_array_equal(rays[:, 7], np.full(n, syn.FAR, np.float32))
    d = fx["rays_d"]
    np.testing.assert_allclose(rays[:, 8:11], d / np.linalg.norm(d, axis=-1, keepdims=True), atol=3e-7)
    # image data: uint8 -> float, background compositing (dataset.py:281-305) -- bit-exact
    np.testing.assert_array_equal(b["target_s"].numpy(), fx["target_s"])
    np.testing.assert_array_equal(b["fgs"].numpy(), fx["fgs"])
    np.testing.assert_array_equal(b["bgs"].numpy(), fx["bgs"])
    # per-ray pose expansion (dataset.py:399-421), image-major (ray_collate_fn)
    np.testing.assert_array_equal(b["kp_batch"].numpy(), fx["kp3d"])
    np.testing.assert_array_equal(b["bones"].numpy(), fx["bones"])
    np.testing.assert_array_equal(b["skts"].numpy(), fx["skts"])
    np.testing.assert_array_equal(b["cyls"].numpy(), fx["cyls"])
    assert b["N_uniques"] == 4
    img, pix = feed.last_idxs
    np.testing.assert_array_equal(img.numpy(), fx["image_idxs"])
    np.testing.assert_array_equal(pix.numpy(), fx["pixel_idxs"])


def test_cam_idxs_default_is_the_queried_index():
    """dataset.py:432-438: the camera code index is the queried image index unless a subclass maps it."""
    fx = np.load(os.path.join(GOLD, "feed_plain.npz"))
    feed, _ = _feed_for(fx)
    b = feed.next_batch(image_idxs=fx["image_idxs"], pixel_idxs=fx["pixel_idxs"])
    assert b["cams"].shape == (48, 1) and b["cams"].dtype == torch.int64
    np.testing.assert_array_equal(b["cams"][:, 0].numpy(), fx["cam_idxs"])


def test_sample_pixels_properties():
    """dataset.py:307-356: N_rand pixels per image, without replacement, inside the sampling mask, increasing; the whole
    image when the mask holds fewer pixels than requested; every mask pixel equally likely."""
    feed = fd.synthetic_feed(n_images=8, H=16, W=16, N_rand=8 * 20, N_sample_images=8)
    g = torch.Generator().manual_seed(0)
    idx = torch.arange(8)
    mask = feed.sampling_masks > 0
    counts = torch.zeros(16 * 16)
    reps = 400
    for _ in range(reps):
        pix = feed.sample_pixels(idx, g)
        assert pix.shape == (8, 20)
        assert (pix[:, 1:] > pix[:, :-1]).all()                       # increasing => distinct
        assert torch.gather(mask, 1, pix).all()
        counts += torch.bincount(pix.reshape(-1), minlength=256)
    m = mask[0]
    assert counts[~m].sum() == 0
    expect = reps * 8 * 20 / int(m.sum())
    assert (counts[m] - expect).abs().max() < 6 * expect ** 0.5       # ~binomial spread
    # mask smaller than the request -> whole image (dataset.py:318-319)
    feed.sampling_masks[3] = 0
    feed.sampling_masks[3, :5] = 1
    seen_outside = False
    for _ in range(20):
        pix = feed.sample_pixels(torch.tensor([3]), g)
        assert (pix[:, 1:] > pix[:, :-1]).all()
        seen_outside |= bool((pix >= 5).any())
    assert seen_outside


def test_next_batch_draws_distinct_sorted_images_and_perturbs_background_only():
    feed = fd.synthetic_feed(n_images=12, H=16, W=16, N_rand=6 * 10, N_sample_images=6, perturb_bg=True, mask_img=True)
    g = torch.Generator().manual_seed(1)
    seen = set()
    for _ in range(30):
        b = feed.next_batch(g)
        img, pix = feed.last_idxs
        assert img.shape == (6,) and (img[1:] > img[:-1]).all()       # RayImageSampler: distinct, np.sort (dataset.py:959-975)
        seen.update(img.tolist())
        fg = b["fgs"]
        plain = torch.gather(feed.bkgds[feed.bkgd_idxs[img]], 1, pix[..., None].expand(-1, -1, 3)).float().reshape(-1, 3) / 255.
        inside = fg[:, 0] > 0
        assert torch.equal(b["bgs"][inside], plain[inside])            # foreground keeps the stored background
        if (~inside).any():
            assert not torch.equal(b["bgs"][~inside], plain[~inside])  # the rest is noise in [0, 1)
            assert (b["bgs"][~inside] >= 0).all() and (b["bgs"][~inside] < 1).all()
            raw = torch.gather(feed.imgs[img], 1, pix[..., None].expand(-1, -1, 3)).float().reshape(-1, 3) / 255.
            assert torch.equal(b["target_s"][~inside], b["bgs"][~inside])      # mask_img: background pixels show bg
            assert torch.equal(b["target_s"][inside], raw[inside])
        assert b["ray_batch"].shape == (60, 11) and b["ray_batch"].is_contiguous()
        assert b["skts"].shape == (60, 24, 4, 4) and b["kp_batch"].shape == (60, 24, 3)
        assert b["N_uniques"] == 6
        # the pose tensors are per-image expands: ray r belongs to pose r // rays_per_image (encoders.py:465-471)
        assert torch.equal(b["skts"][::10], feed.skts[img])
    assert len(seen) == 12


def test_unsupported_flags_raise():
    with pytest.raises(NotImplementedError):
        fd.synthetic_feed(n_images=2, H=8, W=8, patch_size=2)
    with pytest.raises(NotImplementedError):
        fd.synthetic_feed(n_images=2, H=8, W=8, N_nms=1)


@pytest.mark.skipif(not os.path.isdir("/root/reference/core"), reason="authoring container only")
def test_live_reference_other_draws():
    """The unmodified reference data pipeline, live, on draws that are not among the fixtures (focal per axis,
    identity camera rotation short-cut of dataset.py:393-394, no background arrays)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gen_golden_feed as gg
    arrays = fd.synthetic_arrays(n_images=5, H=18, W=14, seed=9, centers=True)
    arrays["focals"] = np.stack([arrays["focals"], arrays["focals"] * 1.1], -1)     # (N, 2): fx, fy
    arrays["c2ws"][1, :3, :3] = np.eye(3, dtype=np.float32)
    for drop_bg in (False, True):
        a = {k: v for k, v in arrays.items() if not (drop_bg and k.startswith("bkgd"))}
        ref = gg.run_reference(a, np.array([1, 3, 4]), rays_per_image=9, seed=5)
        feed = fd.RayFeed.from_arrays(a, 1.0, 5.0, N_rand=27, N_sample_images=3, perturb_bg=False)
        b = feed.next_batch(image_idxs=ref["image_idxs"], pixel_idxs=ref["pixel_idxs"])
        np.testing.assert_array_equal(b["ray_batch"][:, :3].numpy(), ref["rays_o"])
        np.testing.assert_allclose(b["ray_batch"][:, 3:6].numpy(), ref["rays_d"], rtol=0, atol=2e-7)
        np.testing.assert_array_equal(b["target_s"].numpy(), ref["target_s"])
        np.testing.assert_array_equal(b["fgs"].numpy(), ref["fgs"])
        assert ("bgs" in b) == (not drop_bg)
        if not drop_bg:
            np.testing.assert_array_equal(b["bgs"].numpy(), ref["bgs"])
