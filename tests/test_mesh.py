"""Marching cubes + writers (`mesh.py`).  CPU: the extraction is torch ops, the same code runs on the
lattice's device.  mcubes / trimesh are absent (un-vendored, unpinned), so the triangle lists are unpinned; pinned here:
the vertex set (every sign-changing lattice edge, linearly interpolated - mcubes' definition), closed consistently
oriented surfaces also on fields full of ambiguous faces, Euler characteristics, enclosed volumes, file round trips."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import danbo_b200                                    # noqa: E402,F401
from danbo_b200 import mesh                          # noqa: E402


def grid(n, lo=-1., hi=1.):
    g = torch.linspace(lo, hi, n)
    return torch.meshgrid(g, g, g, indexing="ij")


def check_closed_oriented(t):
    """Every undirected edge is shared by exactly two triangles, traversed once in each direction."""
    e = torch.cat([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], 0)
    key = e[:, 0] * (int(t.max()) + 1) + e[:, 1]
    rev = e[:, 1] * (int(t.max()) + 1) + e[:, 0]
    uk, counts = torch.unique(key, return_counts=True)
    assert int(counts.max()) == 1, "a directed edge is used twice: inconsistent orientation or a fin"
    assert torch.equal(uk, torch.unique(rev)), "an edge has no opposite partner: the surface has a hole"
    assert (t[:, 0] != t[:, 1]).all() and (t[:, 1] != t[:, 2]).all() and (t[:, 0] != t[:, 2]).all()
    return e.shape[0] // 2


def test_tables():
    n_tri, tri = mesh.tables()
    assert n_tri.shape == (256,) and int(n_tri[0]) == 0 and int(n_tri[255]) == 0 and int(n_tri.max()) == 5
    assert all(int(n_tri[1 << c]) == 1 for c in range(8))                  # one inside corner: one triangle
    used = (tri >= 0).all(-1).sum(-1)
    assert torch.equal(used, n_tri)


def test_sphere_volume_area_and_topology():
    x, y, z = grid(48)
    r = 0.7
    vol = r - torch.sqrt(x * x + y * y + z * z)                            # inside = positive
    v, t = mesh.marching_cubes(vol, 0.0)
    E = check_closed_oriented(t)
    assert v.shape[0] - E + t.shape[0] == 2                                # Euler characteristic of a sphere
    h = 2.0 / 47
    got = mesh.mesh_volume(v, t) * h ** 3
    assert got > 0 and abs(got - 4 / 3 * np.pi * r ** 3) < 0.01 * 4 / 3 * np.pi * r ** 3
    p = v * h - 1.0
    assert float((p.norm(dim=-1) - r).abs().max()) < 0.3 * h              # vertices lie on the sphere (interpolation error)


def test_torus_and_two_components():
    x, y, z = grid(56)
    q = torch.sqrt(x * x + y * y) - 0.6
    torus = 0.2 - torch.sqrt(q * q + z * z)
    v, t = mesh.marching_cubes(torus, 0.0)
    E = check_closed_oriented(t)
    assert v.shape[0] - E + t.shape[0] == 0                                # genus 1
    two = torch.maximum(0.3 - torch.sqrt((x - 0.5) ** 2 + y * y + z * z), 0.25 - torch.sqrt((x + 0.5) ** 2 + y * y + z * z))
    v, t = mesh.marching_cubes(two, 0.0)
    E = check_closed_oriented(t)
    assert v.shape[0] - E + t.shape[0] == 4                                # two spheres
    h = 2.0 / 55
    want = 4 / 3 * np.pi * (0.3 ** 3 + 0.25 ** 3)
    assert abs(mesh.mesh_volume(v, t) * h ** 3 - want) < 0.02 * want


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_noise_field_is_watertight_and_vertex_set_is_the_edge_crossings(seed):
    """White noise: almost every cell is a surface cell and ambiguous faces abound - the face rule must keep the
    surface closed and consistently oriented.  The vertex set must be exactly mcubes': one vertex per sign-changing
    lattice edge at the linearly interpolated crossing."""
    g = torch.Generator().manual_seed(seed)
    n = 14
    vol = torch.full((n, n, n), -1.0)
    vol[1:-1, 1:-1, 1:-1] = torch.randn(n - 2, n - 2, n - 2, generator=g)  # outside on the lattice boundary: closed surface
    thr = 0.1
    v, t = mesh.marching_cubes(vol, thr)
    check_closed_oriented(t)
    assert int(torch.unique(t).numel()) == v.shape[0]                      # every vertex is used
    # brute-force edge crossings
    want = []
    a = vol.numpy()
    for axis in range(3):
        for i in range(n):
            for j in range(n):
                for k in range(n):
                    p = [i, j, k]
                    if p[axis] + 1 >= n:
                        continue
                    q = list(p); q[axis] += 1
                    v0, v1 = a[tuple(p)], a[tuple(q)]
                    if (v0 > thr) != (v1 > thr):
                        pt = np.array(p, dtype=np.float64)
                        pt[axis] += (thr - v0) / (v1 - v0)
                        want.append(pt)
    want = np.array(want)
    assert want.shape[0] == v.shape[0]
    key = lambda arr: arr[np.lexsort((arr[:, 2], arr[:, 1], arr[:, 0]))]
    np.testing.assert_allclose(key(np.round(v.double().numpy(), 5)), key(np.round(want, 5)), atol=2e-5)
    # the inside volume equals the number of inside lattice points up to the interpolation of the boundary cells
    assert mesh.mesh_volume(v, t) > 0


def test_empty_and_full_volumes():
    v, t = mesh.marching_cubes(torch.zeros(5, 5, 5), 0.5)
    assert v.shape == (0, 3) and t.shape == (0, 3)
    v, t = mesh.marching_cubes(torch.ones(5, 5, 5), 0.5)
    assert v.shape == (0, 3) and t.shape == (0, 3)


def test_ply_and_png_round_trip(tmp_path):
    x, y, z = grid(20)
    v, t = mesh.marching_cubes(0.6 - torch.sqrt(x * x + y * y + z * z), 0.0)
    path = os.path.join(tmp_path, "m.ply")
    mesh.write_ply(path, v / 19 - 0.5, t)
    v2, t2 = mesh.read_ply(path)
    np.testing.assert_array_equal(v2, (v / 19 - 0.5).numpy().astype(np.float32))
    np.testing.assert_array_equal(t2, t.numpy().astype(np.int32))
    head = open(path, "rb").read(200)
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\n") and b"property list uchar int vertex_indices" in head
    from PIL import Image
    rng = np.random.RandomState(0)
    for shape in ((37, 53, 3), (16, 9)):
        img = rng.randint(0, 256, shape).astype(np.uint8)
        p = os.path.join(tmp_path, "a.png")
        mesh.write_png(p, img)
        np.testing.assert_array_equal(np.asarray(Image.open(p)), img)
    with pytest.raises(ValueError):
        mesh.write_png(os.path.join(tmp_path, "b.png"), np.zeros((4, 4, 3), np.float32))


def test_render_mesh_on_a_stand_in_caster(tmp_path):
    """run_render.py:1266-1281 with a caster that returns an analytic lattice (the real `fwd_type='mesh'` needs a GPU)."""
    res = 31

    def caster(kps=None, skts=None, bones=None, radius=1.8, res=31, render_kwargs=None, fwd_type=""):
        assert fwd_type == "mesh" and kps.shape[0] == 1
        x, y, z = grid(res + 1, -radius, radius)
        return 40.0 * (1.0 - torch.sqrt(x * x + y * y + z * z)) - 5.0        # raw density, negative outside
    kps = torch.zeros(2, 24, 3)
    out = mesh.render_mesh(caster, kps, torch.zeros(2, 24, 4, 4), torch.zeros(2, 24, 3), res=res, threshold=10.,
                           out_dir=str(tmp_path))
    assert len(out) == 2 and os.path.exists(os.path.join(tmp_path, "meshes", "001.ply"))
    v, t = out[0]
    check_closed_oriented(t)
    assert float(v.min()) >= -0.5 and float(v.max()) <= 0.5                 # vertices / res - 0.5
    # sigma = 10  <=>  r = 1 - 15/40 = 0.625 world units = 0.625 / 3.6 of the unit cube
    r = ((v * 1.0).norm(dim=-1))
    assert float((r - 0.625 / 3.6).abs().max()) < 0.01


def test_save_renders_writes_pngs_and_video(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from PIL import Image
    rng = np.random.RandomState(1)
    rgbs = torch.tensor(rng.rand(5, 32, 48, 3).astype(np.float32))
    accs = torch.tensor(rng.rand(5, 32, 48).astype(np.float32))
    mesh.save_renders(str(tmp_path), rgbs, accs, fps=10)
    img = np.asarray(Image.open(os.path.join(tmp_path, "image", "00003.png")))
    np.testing.assert_array_equal(img, (rgbs[3].clamp(0, 1) * 255).to(torch.uint8).numpy())
    acc = np.asarray(Image.open(os.path.join(tmp_path, "acc", "00000.png")))
    assert acc.shape == (32, 48)
    cap = cv2.VideoCapture(os.path.join(tmp_path, "render_rgb.mp4"))
    n = 0
    while cap.read()[0]:
        n += 1
    assert n == 5 and abs(cap.get(cv2.CAP_PROP_FPS) - 10) < 0.5
    with pytest.raises(ValueError):
        mesh.write_video(os.path.join(tmp_path, "x.mp4"), np.zeros((2, 4, 4, 3), np.float32))
