"""CPU property tests of the two conservative prefilters the kernels use (csrc/field.cu): they may only ever skip work
whose result is known, never change a result.  The kernels' fp32 arithmetic is restated in numpy float32 and checked
against float64 ground truth on a million random (ray, bone) pairs, including grazing and axis-parallel rays.

  * `mark_candidate` (sample_mask): a bone is dropped from a ray's candidate list only if the affine map x(z) = c + z e
    stays outside |x|_inf <= 1 + 2e-4 (the kernel's decision margin) for EVERY depth z;
  * the near/far line prefilter (nearfar_finish): the fp64 plane intersections are skipped only if none of the six plane
    points can lie inside the box |p|_inf <= hi."""
import numpy as np

F = np.float32


def _random_maps(n, rng):
    """Affine maps of the size the kernels see: box units, origins up to ~60 boxes away, plus degenerate directions."""
    c = (rng.standard_normal((n, 3)) * rng.choice([0.5, 3.0, 30.0], (n, 1))).astype(F)
    e = (rng.standard_normal((n, 3)) * rng.choice([0.2, 2.0, 20.0], (n, 1))).astype(F)
    e[rng.random((n, 3)) < 0.03] = 0.0                                   # axis-parallel rays
    e[rng.random((n, 3)) < 0.02] *= F(1e-9)                              # almost axis-parallel
    # aim a share of the lines at the box so that hits, grazes and near misses are all well represented
    aim = rng.random(n) < 0.6
    target = (rng.uniform(-1.3, 1.3, (n, 3))).astype(F)
    z0 = rng.uniform(1.0, 5.0, n).astype(F)
    c[aim] = (target[aim] - z0[aim, None] * e[aim]).astype(F)
    return c, e


def _candidate_fp32(c, e, cull=F(4e-4)):
    """numpy restatement of `mark_candidate`: True = the bone stays in the ray's candidate list."""
    one = F(1.0)
    lo = np.full(len(c), F(-3.0e38)); hi = np.full(len(c), F(3.0e38))
    empty = np.zeros(len(c), bool)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for i in range(3):
            tiny = np.abs(e[:, i]) < F(1e-20)
            empty |= tiny & (np.abs(c[:, i]) > one + cull)
            inv = (one / e[:, i]).astype(F)
            a = ((-(one + cull) - c[:, i]) * inv).astype(F)
            b = (((one + cull) - c[:, i]) * inv).astype(F)
            lo = np.where(tiny, lo, np.maximum(lo, np.minimum(a, b)))
            hi = np.where(tiny, hi, np.minimum(hi, np.maximum(a, b)))
    return ~(empty | (lo > hi))


def test_candidate_culling_never_drops_a_bone_a_sample_could_need():
    rng = np.random.default_rng(0)
    c, e = _random_maps(1_000_000, rng)
    keep = _candidate_fp32(c, e)
    # ground truth in float64: does |c + z e|_inf <= 1 + 2e-4 hold for some z?  (slab intersection, exact enough in fp64)
    c64, e64 = c.astype(np.float64), e.astype(np.float64)
    m = 1.0 + 2e-4
    lo = np.full(len(c), -np.inf); hi = np.full(len(c), np.inf)
    feasible = np.ones(len(c), bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(3):
            zero = e64[:, i] == 0.0
            feasible &= ~(zero & (np.abs(c64[:, i]) > m))
            a = (-m - c64[:, i]) / e64[:, i]
            b = (m - c64[:, i]) / e64[:, i]
            lo = np.where(zero, lo, np.maximum(lo, np.minimum(a, b)))
            hi = np.where(zero, hi, np.minimum(hi, np.maximum(a, b)))
    needed = feasible & (lo <= hi)
    assert needed.sum() > 100_000 and (~needed).sum() > 100_000          # both outcomes are well represented
    dropped_but_needed = needed & ~keep
    assert not dropped_but_needed.any(), int(dropped_but_needed.sum())
    # and it is a useful filter: most bones that cannot be needed are dropped
    assert (keep & ~needed).sum() < 0.02 * (~needed).sum()


def _nearfar_prefilter_fp32(os_, ds, hi=F(1.3001)):
    H = F(hi * F(1.002) + F(2e-3))
    lo_t = np.full(len(os_), F(-3.0e38)); hi_t = np.full(len(os_), F(3.0e38))
    may = np.ones(len(os_), bool)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for i in range(3):
            tiny = np.abs(ds[:, i]) < F(1e-12)
            may &= ~(tiny & (np.abs(os_[:, i]) > H))
            inv = (F(1.0) / ds[:, i]).astype(F)
            ta = ((-H - os_[:, i]) * inv).astype(F)
            tb = ((H - os_[:, i]) * inv).astype(F)
            lo_t = np.where(tiny, lo_t, np.maximum(lo_t, np.minimum(ta, tb)))
            hi_t = np.where(tiny, hi_t, np.minimum(hi_t, np.maximum(ta, tb)))
        slack = (F(1e-3) * (F(1.0) + np.abs(lo_t) + np.abs(hi_t))).astype(F)
        may &= ~(lo_t > hi_t + slack)
    return may


def test_nearfar_prefilter_never_skips_a_box_with_a_plane_hit():
    """Ground truth = the kernel's own fp64 rule (ray_utils.py:383-417): a plane point counts when all three coordinates
    lie within +-hi; the prefilter may say "skip" only when no plane point counts."""
    rng = np.random.default_rng(1)
    os_, ds = _random_maps(1_000_000, rng)
    may = _nearfar_prefilter_fp32(os_, ds)
    hi, bound = np.float64(F(1.3001)), 1.3
    o64, d64 = os_.astype(np.float64), ds.astype(np.float64)
    any_hit = np.zeros(len(os_), bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in range(6):
            a = k % 3
            b = -bound if k < 3 else bound
            t = (b - o64[:, a]) / d64[:, a]
            p = (t[:, None] * d64 + o64).astype(F).astype(np.float64)    # the kernel rounds the plane point to fp32
            inside = np.all((p <= hi) & (p >= -hi), axis=1)              # NaN / inf compare false, as on the device
            any_hit |= inside
    assert any_hit.sum() > 100_000 and (~any_hit).sum() > 100_000
    assert not (any_hit & ~may).any(), int((any_hit & ~may).sum())
    assert (may & ~any_hit).sum() < 0.05 * (~any_hit).sum()


def test_empty_sample_constants_identity():
    """The identity behind `empty_trunk_kernel` / `empty_rows_kernel` (csrc/mlp_tcgen05.cu): for a sample no bone sees the
    blended feature is 0, the MLP input is PE(0), and the field's output is
        sigma0 = alpha(trunk(PE(0))),   rgb = W_rgb relu(c + ray_bias) + b_rgb,   c = W_v[:, :256] feat(trunk(PE(0))),
    with ray_bias = W_v[:, 256:] [PE(d) ; code] + b_v the only per-ray quantity.  Checked against the oracle's `field_mlp`
    (core/networks/nerf.py:176-209) on random view inputs."""
    import sys, os
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import danbo_oracle as orc
    from danbo_b200 import synthetic as syn
    P = syn.synthetic_params(0)
    torch.manual_seed(0)
    n = 64
    view = torch.randn(n, 155) * 0.5
    x0 = orc.pe_embed(torch.zeros(1, 15), 6)                                  # PE(0): zeros, then (sin 0, cos 0) per octave
    assert x0.shape == (1, 195) and float(x0.sum()) == 6 * 15
    ones_at = [30 + 30 * f + i for f in range(6) for i in range(15)]          # the kernel's own index rule
    assert torch.equal(torch.nonzero(x0[0])[:, 0], torch.tensor(ones_at))
    want = orc.field_mlp(x0.expand(n, -1), view, P)
    h = orc.density_trunk(x0, P)
    sigma0 = h @ P["alpha_linear.weight"].t() + P["alpha_linear.bias"]
    feat = h @ P["feature_linear.weight"].t() + P["feature_linear.bias"]
    Wv = P["views_linears.0.weight"]
    c = feat @ Wv[:, :256].t()
    ray_bias = view @ Wv[:, 256:].t() + P["views_linears.0.bias"]
    rgb = torch.relu(c + ray_bias) @ P["rgb_linear.weight"].t() + P["rgb_linear.bias"]
    got = torch.cat([rgb, sigma0.expand(n, 1)], -1)
    assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())
