"""Training data feed for the path's trainer: the whole training set resident on the device and ray
sampling done there, replacing the h5py `DataLoader` of the reference at the point where its output enters
`Trainer.train_batch`.

Mirrors core/dataset.py for the flags the shipped configs use (patch_size = 1, N_nms = 0): `BaseH5Dataset.init_meta`
(:148-200, the pre-computed pixel directions), `sample_pixels` (:307-356: N_rand / N_sample_images pixels per image, drawn
without replacement from the image's sampling mask, sorted), `get_rays` (:383-401), `get_img_data` (:281-305: /255,
background compositing with `mask_img`, `perturb_bg`), `RayImageSampler` (:941-976: N_sample_images distinct images per
iteration, sorted) and `ray_collate_fn` (:980-987: image-major flattening).  The arrays have the layout of the reference's
HDF5 keys (`imgs, masks, sampling_masks` (N, H*W, C) uint8, `kp3d, bones, skts, cyls, c2ws, focals, centers, bkgds,
bkgd_idxs`, `img_shape`), so a converted data set drops in; reading the HDF5 file itself stays with the caller (h5py).

`next_batch()` returns what `training.TrainStep` consumes: `ray_batch` (n, 11) = [o, d, near, far, viewdir] as
`render()` assembles it (core/trainer.py:96-162), `kp_batch, skts, bones, cyls` expanded per ray, `cams`, `target_s`,
`bgs`, `fgs`, `N_uniques`.  Everything is torch ops on the arrays' device: at > 700 iterations/s a host loader that
gathers 16 images from HDF5 per step would be the bottleneck.
"""
import torch


class RayFeed:
    def __init__(self, imgs, masks, sampling_masks, kp3d, bones, skts, cyls, c2ws, focals, img_shape, near, far,
                 centers=None, bkgds=None, bkgd_idxs=None, cam_idxs=None, N_rand=3072, N_sample_images=16,
                 mask_img=True, perturb_bg=True, patch_size=1, N_nms=0, device=None, rank=0, world_size=1, seed=0):
        if patch_size != 1 or N_nms != 0:
            raise NotImplementedError("patch_size > 1 / N_nms > 0 are not implemented (no shipped config sets them)")
        dev = torch.device(device) if device is not None else torch.as_tensor(imgs).device
        t = lambda a, dt=None: None if a is None else torch.as_tensor(a).to(dev, dtype=dt)
        self.device = dev
        self.H, self.W = int(img_shape[-3]) if len(img_shape) == 4 else int(img_shape[0]), \
            int(img_shape[-2]) if len(img_shape) == 4 else int(img_shape[1])
        HW = self.H * self.W
        self.imgs = t(imgs, torch.uint8).reshape(-1, HW, 3)
        self.masks = t(masks, torch.uint8).reshape(-1, HW, 1)
        self.sampling_masks = t(sampling_masks, torch.uint8).reshape(-1, HW)
        self.n_images = self.imgs.shape[0]
        self.kp3d, self.bones, self.skts, self.cyls = (t(a, torch.float32) for a in (kp3d, bones, skts, cyls))
        self.c2ws, self.focals = t(c2ws, torch.float32), t(focals, torch.float32)
        self.centers = t(centers, torch.float32)
        self.bkgds = None if bkgds is None else t(bkgds, torch.uint8).reshape(-1, HW, 3)
        self.bkgd_idxs = t(bkgd_idxs, torch.int64)
        self.cam_idxs = torch.arange(self.n_images, device=dev) if cam_idxs is None else t(cam_idxs, torch.int64)
        self.near, self.far = float(near), float(far)
        self.N_sample_images = int(N_sample_images)
        self.rays_per_image = int(N_rand) // int(N_sample_images)
        # data parallel: every rank draws the SAME images (a host generator seeded alike everywhere) and
        # keeps a contiguous run of the sorted batch, so the ranks' batches concatenated are the global image-major batch
        self.rank, self.world_size = int(rank), int(world_size)
        if self.N_sample_images % self.world_size:
            raise ValueError(f"N_sample_images={N_sample_images} is not a multiple of world_size={world_size}")
        self._image_gen = torch.Generator().manual_seed(int(seed))
        self._perm, self._cursor = None, 0
        self.mask_img, self.perturb_bg = bool(mask_img), bool(perturb_bg)
        # dataset.py:163-183: pixel directions before the division by the focal length (x right, y up, looking down -z)
        j, i = torch.meshgrid(torch.arange(self.H, dtype=torch.float32, device=dev),
                              torch.arange(self.W, dtype=torch.float32, device=dev), indexing="ij")
        i, j = i.reshape(-1), j.reshape(-1)
        if self.centers is None:
            off_y, off_x = self.H * 0.5, self.W * 0.5
        else:
            off_y = off_x = 0.
        self._dirs = torch.stack([i - off_x, -(j - off_y), -torch.ones_like(i)], -1)
        self.rebuild_sampling_index()

    # ---- dataset.py:307-356 -------------------------------------------------------------------------------------
    def rebuild_sampling_index(self):
        """Compressed list of every image's candidate pixels (CSR: `_valid_flat[_valid_off[i] : _valid_off[i+1]]`,
        increasing): the sampling mask, or the whole image when the mask holds fewer pixels than one draw needs
        (dataset.py:317-319).  Call again after editing `sampling_masks`."""
        counts, flat = [], []
        for s in range(0, self.n_images, 64):                               # 64 images at a time bounds the temporaries
            m = self.sampling_masks[s:s + 64] > 0
            m = m | (m.sum(-1, keepdim=True) < self.rays_per_image)
            counts.append(m.sum(-1))
            flat.append(torch.nonzero(m)[:, 1].to(torch.int32))             # row-major: per image, increasing
        self._n_valid = torch.cat(counts) if counts else torch.zeros(0, dtype=torch.int64, device=self.device)
        off = torch.zeros(self.n_images + 1, dtype=torch.int64, device=self.device)
        off[1:] = torch.cumsum(self._n_valid, 0)
        self._valid_off = off
        self._valid_flat = torch.cat(flat) if flat else torch.zeros(0, dtype=torch.int32, device=self.device)
        # rejection-free fast path needs collisions among the candidates to be rare (see sample_pixels)
        self._sparse_draw = bool((self._n_valid >= 64 * self.rays_per_image).all()) if self.n_images else False

    def sample_pixels(self, image_idxs, generator=None):
        """(B,) image indices -> (B, rays_per_image) increasing flat pixel indices, uniform without replacement over each
        image's candidate pixels.  Two exact samplers, chosen once per data set (no per-step host decision):

        * masks of >= 64 R pixels (real data: ~1e5 foreground pixels, R = 192): draw 2R candidates WITH replacement and
          keep the first R distinct ones in draw order - sequentially skipping repeats is a uniform draw without
          replacement; fewer than R distinct among 2R has probability < (2R choose R)·(R/n)^R, below 1e-100 here.
          Cost O(B·R), independent of the image size;
        * otherwise random keys over the image's pixels and the R largest (cost O(B·H·W))."""
        R = self.rays_per_image
        B = image_idxs.shape[0]
        if self._sparse_draw:
            n = self._n_valid[image_idxs][:, None]
            u = torch.rand(B, 2 * R, device=self.device, generator=generator, dtype=torch.float64)
            cand = torch.minimum((u * n).long(), n - 1)
            val, perm = torch.sort(cand, dim=-1, stable=True)
            rep_sorted = torch.zeros_like(cand)
            rep_sorted[:, 1:] = (val[:, 1:] == val[:, :-1]).long()           # stable sort: the later draw is the repeat
            rep = torch.zeros_like(cand).scatter_(1, perm, rep_sorted)
            order = torch.arange(2 * R, device=self.device)[None] + rep * (2 * R)
            first = torch.topk(order, R, dim=-1, largest=False).indices      # the first R distinct draws
            k = torch.sort(torch.gather(cand, 1, first), -1).values
            return self._valid_flat[self._valid_off[image_idxs][:, None] + k].long()
        m = self.sampling_masks[image_idxs] > 0
        m = m | (m.sum(-1, keepdim=True) < R)
        keys = torch.rand(m.shape, device=self.device, generator=generator)
        keys = torch.where(m, keys, torch.full_like(keys, -1.))
        pix = torch.topk(keys, R, dim=-1).indices                            # random keys: a uniform draw without replacement
        return torch.sort(pix, -1).values

    # ---- dataset.py:915-976 -------------------------------------------------------------------------------------
    def draw_images(self):
        """The next N_sample_images image indices as `RayImageSampler` yields them: consecutive entries of a random
        permutation of the data set (every image once per pass; a batch that straddles two passes continues into a
        fresh permutation), sorted.  Host-side and seeded alike on every rank; this rank's share is returned."""
        batch = []
        while len(batch) < self.N_sample_images:
            if self._perm is None or self._cursor >= self.n_images:
                self._perm, self._cursor = torch.randperm(self.n_images, generator=self._image_gen), 0
            take = min(self.N_sample_images - len(batch), self.n_images - self._cursor)
            batch += self._perm[self._cursor:self._cursor + take].tolist()
            self._cursor += take
        full = torch.sort(torch.tensor(batch, dtype=torch.int64)).values
        per = self.N_sample_images // self.world_size
        return full[self.rank * per:(self.rank + 1) * per]

    # ---- dataset.py:383-401 -------------------------------------------------------------------------------------
    def get_rays(self, image_idxs, pix):
        dirs = self._dirs[pix]                                               # (B, R, 3)
        if self.centers is not None:
            c = self.centers[image_idxs].clone()
            c[:, 1] = -c[:, 1]
            dirs = torch.cat([dirs[..., :2] - c[:, None, :], dirs[..., 2:]], -1)
        focal = self.focals[image_idxs].reshape(len(image_idxs), 1, -1)
        dirs = torch.cat([dirs[..., :2] / focal, dirs[..., 2:]], -1)
        c2w = self.c2ws[image_idxs]
        rays_d = torch.sum(dirs[..., None, :] * c2w[:, None, :3, :3], -1)
        rays_o = c2w[:, None, :3, 3].expand_as(rays_d)
        return rays_o, rays_d

    # ---- dataset.py:281-305 -------------------------------------------------------------------------------------
    def get_img_data(self, image_idxs, pix, generator=None):
        # only the sampled pixels are touched: (image, pixel) -> row of the flattened (N*H*W, C) array
        HW = self.H * self.W
        gather = lambda a, rows: a.reshape(-1, a.shape[-1])[rows[:, None] * HW + pix]
        fg = gather(self.masks, image_idxs).float()
        # a TENSOR divisor: torch's CUDA kernels turn division by a Python scalar into a multiplication by its reciprocal,
        # which is 1 ulp off the reference's true division for about half of the 256 byte values
        s255 = torch.full((), 255., device=self.device)
        img = gather(self.imgs, image_idxs).float() / s255
        bg = None
        if self.bkgds is not None:
            bg = gather(self.bkgds, self.bkgd_idxs[image_idxs]).float() / s255
            if self.perturb_bg:
                noise = torch.rand(bg.shape, device=self.device, generator=generator)
                bg = (1 - fg) * noise + fg * bg
            if self.mask_img:
                img = img * fg + (1. - fg) * bg
        return img, fg, bg

    # ---- dataset.py:941-987 + trainer.py:96-162 -----------------------------------------------------------------
    def next_batch(self, generator=None, image_idxs=None, pixel_idxs=None):
        """One training batch.  `image_idxs` / `pixel_idxs` ((B,) and (B, R), both increasing) replace the random draws
        (that is how the parity test replays the reference's draws); the indices used are kept in `self.last_idxs`."""
        R = self.rays_per_image
        if image_idxs is None:
            image_idxs = self.draw_images()
        image_idxs = torch.sort(torch.as_tensor(image_idxs).long()).values.to(self.device, non_blocking=True)
        B = image_idxs.shape[0]
        if pixel_idxs is None:
            pix = self.sample_pixels(image_idxs, generator)
        else:
            pix = torch.as_tensor(pixel_idxs, device=self.device).long().reshape(B, -1)
            R = pix.shape[1]
        rays_o, rays_d = self.get_rays(image_idxs, pix)
        img, fg, bg = self.get_img_data(image_idxs, pix, generator)
        n = B * R
        flat = lambda a: a.reshape(n, *a.shape[2:])
        rays_o, rays_d = flat(rays_o), flat(rays_d)
        ones = torch.ones(n, 1, device=self.device)
        view = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
        per_ray = lambda a: a[image_idxs][:, None].expand(B, R, *a.shape[1:]).reshape(n, *a.shape[1:])
        batch = {"ray_batch": torch.cat([rays_o, rays_d, self.near * ones, self.far * ones, view], -1).contiguous(),
                 "kp_batch": per_ray(self.kp3d), "skts": per_ray(self.skts), "bones": per_ray(self.bones),
                 "cyls": per_ray(self.cyls), "cams": per_ray(self.cam_idxs[:, None]), "target_s": flat(img),
                 "fgs": flat(fg), "kp_idx": image_idxs[:, None].expand(B, R).reshape(n), "N_uniques": B}
        if bg is not None:
            batch["bgs"] = flat(bg)
        self.last_idxs = (image_idxs, pix)
        return batch

    def prefetch(self, generator=None):
        """`next_batch` one iteration ahead on a side stream (CUDA arrays only; plain `next_batch` otherwise): the ~25
        small launches of a draw overlap the previous training iteration instead of preceding the next one - the role
        the reference gives its DataLoader workers (run_nerf.py:596-600).  The returned batch is safe to use on the
        caller's current stream."""
        if self.device.type != "cuda":
            return self.next_batch(generator)
        cur = torch.cuda.current_stream(self.device)
        if getattr(self, "_side", None) is None:
            self._side, self._pending = torch.cuda.Stream(device=self.device), None

        def produce():
            self._side.wait_stream(cur)                       # the arrays (and a generator's state) as the caller left them
            with torch.cuda.stream(self._side):
                b = self.next_batch(generator)
                ev = torch.cuda.Event()
                ev.record(self._side)
            return b, ev, self.last_idxs
        if self._pending is None:
            self._pending = produce()
        batch, ev, idxs = self._pending
        cur.wait_event(ev)
        for v in batch.values():
            if torch.is_tensor(v):
                v.record_stream(cur)                          # allocated on the side stream, consumed on the caller's
        self._pending = produce()
        self.last_idxs = idxs
        return batch

    # ---- the reference's HDF5 keys (dataset.py:155-205, process_spin.py:234-297) ---------------------------------
    @classmethod
    def from_arrays(cls, d, near, far, **kw):
        """`d`: mapping with the reference's HDF5 keys (an open h5py.File, an np.load()ed archive or a dict of arrays).
        `centers`, `bkgds` / `bkgd_idxs` are optional, as in the reference (`'centers' in dataset`, `has_bg`)."""
        opt = lambda k: d[k][:] if k in d else None
        return cls(imgs=d["imgs"][:], masks=d["masks"][:], sampling_masks=d["sampling_masks"][:], kp3d=d["kp3d"][:],
                   bones=d["bones"][:], skts=d["skts"][:], cyls=d["cyls"][:], c2ws=d["c2ws"][:], focals=d["focals"][:],
                   img_shape=d["img_shape"][:], near=near, far=far, centers=opt("centers"), bkgds=opt("bkgds"),
                   bkgd_idxs=opt("bkgd_idxs"), **kw)


def synthetic_arrays(n_images=32, H=128, W=128, seed=0, centers=False):
    """A synthetic training set under the reference's HDF5 keys (random images, a disc as mask, the synthetic poses and
    cameras of SURVEY §8d): what the feed is exercised and benchmarked with, in place of the licensed data."""
    import numpy as np
    from . import synthetic as syn
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    disc = (((yy - H / 2) ** 2 + (xx - W / 2) ** 2) < (0.35 * H) ** 2).astype(np.uint8).reshape(1, H * W, 1)
    poses = [syn.make_pose(seed * n_images + i, render_cylinder=False) for i in range(n_images)]
    cams = syn.bullet_time_cameras(syn.camera(), n_images)
    stack = lambda k: np.stack([np.asarray(p[k], dtype=np.float32) for p in poses])
    d = {"imgs": rng.randint(0, 256, (n_images, H * W, 3)).astype(np.uint8),
         "masks": np.repeat(disc, n_images, 0), "sampling_masks": np.repeat(disc, n_images, 0),
         "kp3d": stack("kps"), "bones": stack("bones"), "skts": stack("skts"), "cyls": stack("cyl"),
         "c2ws": np.stack([np.asarray(c, dtype=np.float32) for c in cams]),
         "focals": np.full((n_images,), 1.2 * H, dtype=np.float32),
         "img_shape": np.array([n_images, H, W, 3], dtype=np.int64),
         "bkgds": rng.randint(0, 256, (2, H * W, 3)).astype(np.uint8),
         "bkgd_idxs": rng.randint(0, 2, (n_images,)).astype(np.int64)}
    if centers:
        d["centers"] = (np.array([W, H], dtype=np.float32) * 0.5 + rng.randn(n_images, 2).astype(np.float32) * 3.0)
    return d


def synthetic_feed(n_images=32, H=128, W=128, device="cpu", seed=0, centers=False, **kw):
    import numpy as np
    from . import synthetic as syn
    return RayFeed.from_arrays(synthetic_arrays(n_images, H, W, seed, centers), syn.NEAR, syn.FAR,
                               cam_idxs=np.arange(n_images) % 8, device=device, seed=seed, **kw)
