"""Synthetic, procedurally textured multi-view scenes with known geometry (no datasets are available offline).

Shapes follow BASELINE.json's configs / SURVEY.md section 8(d): planar facets rendered by analytic ray-plane
intersection, value-noise textures defined on the facets (so every view sees the same surface), cameras sharing K,
`pair.txt`-style neighbour lists, and the reference's on-disk layout (images/%08d.png, cams/%08d_cam.txt, pair.txt;
formats: tools/colmap2mvsnet.py:494-514, ReadCamera APD.cpp:85-135, GenerateSampleList main.cpp:44-102).
"""
import os

import numpy as np

from .binding import Camera


class Facet:
    """Planar quad: points o + a*u + b*v with a in [0, la], b in [0, lb]; texture amplitude amp (grey levels)."""

    def __init__(self, o, u, v, la, lb, amp=100.0, freq=6.0, tex_seed=0, weak_blobs=0.0):
        self.o = np.asarray(o, np.float64)
        self.u = np.asarray(u, np.float64) / np.linalg.norm(u)
        self.v = np.asarray(v, np.float64) / np.linalg.norm(v)
        self.n = np.cross(self.u, self.v)
        self.n /= np.linalg.norm(self.n)
        self.la, self.lb = float(la), float(lb)
        self.amp, self.freq, self.tex_seed, self.weak_blobs = amp, freq, tex_seed, weak_blobs


def _lattice(seed, n=256):
    return np.random.RandomState(seed).rand(n, n).astype(np.float32)


def _value_noise(a, b, lat):
    n = lat.shape[0]
    ia, ib = np.floor(a).astype(np.int64), np.floor(b).astype(np.int64)
    fa, fb = (a - ia).astype(np.float32), (b - ib).astype(np.float32)
    fa = fa * fa * (3 - 2 * fa)
    fb = fb * fb * (3 - 2 * fb)
    i0, i1, j0, j1 = ia % n, (ia + 1) % n, ib % n, (ib + 1) % n
    return (lat[j0, i0] * (1 - fa) + lat[j0, i1] * fa) * (1 - fb) + (lat[j1, i0] * (1 - fa) + lat[j1, i1] * fa) * fb


def _texture(facet, a, b, octaves=5):
    """5-octave value noise around 128; optional smooth blobs where the amplitude drops to <= 1 grey level."""
    acc = np.zeros(a.shape, np.float32)
    norm = 0.0
    for o in range(octaves):
        lat = _lattice(1000 * facet.tex_seed + o)
        f = facet.freq * (2 ** o)
        w = 0.5 ** (o * 0.6)
        acc += w * (_value_noise(a * f, b * f, lat) - 0.5)
        norm += w * 0.5
    tex = acc / norm  # ~[-1, 1]
    amp = np.full(a.shape, facet.amp, np.float32)
    if facet.weak_blobs > 0:
        blob = _value_noise(a * 0.35 + 17.0, b * 0.35 + 5.0, _lattice(7777 + facet.tex_seed))
        amp = np.where(blob < facet.weak_blobs, 1.0, amp).astype(np.float32)
    return 128.0 + amp * tex


def look_at(eye, target, down=(0, 1, 0)):
    """world->camera rotation R and translation t (x_cam = R X + t); camera x right, y down, z forward."""
    eye, target, down = np.asarray(eye, np.float64), np.asarray(target, np.float64), np.asarray(down, np.float64)
    z = target - eye
    z /= np.linalg.norm(z)
    x = np.cross(down, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])
    return R, -R @ eye


def render_view(facets, K, R, t, width, height, chunk=256):
    """returns (gray uint8 HxW, depth float32 HxW camera-z, 0 where nothing is hit)"""
    Kinv = np.linalg.inv(K)
    C = -R.T @ t
    img = np.zeros((height, width), np.float32)
    depth = np.zeros((height, width), np.float32)
    xs = np.arange(width, dtype=np.float64)
    for y0 in range(0, height, chunk):
        y1 = min(height, y0 + chunk)
        ys = np.arange(y0, y1, dtype=np.float64)
        gx, gy = np.meshgrid(xs, ys)
        pix = np.stack([gx, gy, np.ones_like(gx)], -1)
        d = (pix @ Kinv.T) @ R  # rows: R^T K^-1 p, camera z component == 1
        best = np.full(gx.shape, np.inf)
        val = np.full(gx.shape, 128.0, np.float32)
        for f in facets:
            denom = d @ f.n
            num = float(np.dot(f.n, f.o - C))
            with np.errstate(divide="ignore", invalid="ignore"):
                tt = num / denom
            X = C + d * tt[..., None]
            rel = X - f.o
            a, b = rel @ f.u, rel @ f.v
            ok = (tt > 1e-6) & (a >= 0) & (a <= f.la) & (b >= 0) & (b <= f.lb) & (tt < best)
            if ok.any():
                tex = _texture(f, a[ok], b[ok])
                val[ok] = tex
                best[ok] = tt[ok]
        hit = np.isfinite(best)
        img[y0:y1] = val
        depth[y0:y1] = np.where(hit, best, 0.0)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8), depth


class Scene:
    def __init__(self, width, height):
        self.width, self.height = width, height
        self.images, self.colors, self.cameras, self.pairs, self.gt_depth = [], None, [], [], []
        self.K, self.Rs, self.ts = None, [], []

    # ---- reference on-disk layout
    def write_dense_folder(self, folder):
        import cv2
        os.makedirs(os.path.join(folder, "images"), exist_ok=True)
        os.makedirs(os.path.join(folder, "cams"), exist_ok=True)
        for i, img in enumerate(self.images):
            cv2.imwrite(os.path.join(folder, "images", "%08d.png" % i), img)
            cam = self.cameras[i]
            with open(os.path.join(folder, "cams", "%08d_cam.txt" % i), "w") as f:
                f.write("extrinsic\n")
                R, t = self.Rs[i], self.ts[i]
                for r in range(3):
                    f.write("%.9g %.9g %.9g %.9g\n" % (R[r, 0], R[r, 1], R[r, 2], t[r]))
                f.write("0.0 0.0 0.0 1.0\n\nintrinsic\n")
                for r in range(3):
                    f.write("%.9g %.9g %.9g\n" % tuple(self.K[r]))
                f.write("\n%.9g %.9g %d %.9g\n" % (cam.depth_min, cam.interval, int(cam.depth_num), cam.depth_max))
        with open(os.path.join(folder, "pair.txt"), "w") as f:
            f.write("%d\n" % len(self.images))
            for i, src in enumerate(self.pairs):
                f.write("%d\n%d %s\n" % (i, len(src), " ".join("%d 1.0" % s for s in src)))


def _render_job(job):
    facets, K, R, t, W, H = job
    with np.errstate(all="ignore"):
        return render_view(facets, K, R, t, W, H)


def render_views(facets, K, poses, width, height, which=None, workers=0):
    """(gray, depth) of the views `which` (default all), rendered by a pool of processes when there is more than one big view
    (ray casting + 5-octave value noise is ~4 s of one core per 1080p view)"""
    which = list(range(len(poses))) if which is None else list(which)
    jobs = [(facets, K, poses[i][0], poses[i][1], width, height) for i in which]
    workers = workers or min(len(jobs), os.cpu_count() or 1, 16)
    if workers <= 1 or len(jobs) < 2 or width * height < 256 * 256:
        return [_render_job(j) for j in jobs]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(_render_job, jobs, chunksize=1)


def _finish(scene, facets, K, poses, num_src, depth_range, with_color=False, prerendered=None):
    W, H = scene.width, scene.height
    scene.K = K
    centers = []
    if prerendered is None:
        prerendered = render_views(facets, K, poses, W, H)
    for i, (R, t) in enumerate(poses):
        gray, depth = prerendered[i]
        scene.images.append(gray)
        scene.gt_depth.append(depth)
        scene.Rs.append(R)
        scene.ts.append(t)
        scene.cameras.append(Camera.make(K, R, t, W, H, depth_range[0], depth_range[1]))
        centers.append(-R.T @ t)
    if with_color:
        scene.colors = [np.stack([g, np.roll(g, 1, 0), 255 - g], -1) for g in scene.images]
    centers = np.array(centers)
    for i in range(len(poses)):
        dist = np.linalg.norm(centers - centers[i], axis=1)
        dist[i] = np.inf
        order = np.argsort(dist, kind="stable")
        scene.pairs.append([int(j) for j in order[:min(num_src, len(poses) - 1)]])
    return scene


def make_plane_scene(width=640, height=480, num_views=5, num_src=4, seed=1, with_color=False):
    """C1: one slanted textured plane z = 4 + 0.15x - 0.1y, cameras on a 0.8 m line looking at the plane centre."""
    f = 0.9 * width
    K = np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1.0]])
    nrm = np.array([-0.15, 0.1, 1.0])
    o = np.array([-6.0, -6.0, 4 + 0.15 * -6.0 - 0.1 * -6.0])
    u = np.array([1.0, 0, 0.15])
    v = np.cross(nrm, u)
    facets = [Facet(o, u, v, 13.0, 13.0, amp=100.0, freq=5.0, tex_seed=seed)]
    poses = []
    for i in range(num_views):
        x = -0.4 + 0.8 * i / max(1, num_views - 1)
        poses.append(look_at([x, 0.02 * i, 0.0], [0.0, 0.0, 4.0]))
    return _finish(Scene(width, height), facets, K, poses, num_src, (2.0, 8.0), with_color)


def _room_facets(seed, weak=0.0):
    fl = Facet([-6, 1.5, 0], [1, 0, 0], [0, 0, 1], 12, 12, amp=90, freq=4.0, tex_seed=seed, weak_blobs=weak)
    back = Facet([-6, -4.5, 8], [1, 0, 0], [0, 1, 0], 12, 6, amp=100, freq=4.5, tex_seed=seed + 1, weak_blobs=weak)
    left = Facet([-5, -4.5, 0], [0, 0, 1], [0, 1, 0], 8, 6, amp=100, freq=4.0, tex_seed=seed + 2, weak_blobs=weak)
    fac = [fl, back, left]
    # three boxes: front + top + one side each
    for k, (bx, bz, s, hgt) in enumerate([(-2.5, 5.0, 1.2, 1.4), (0.3, 6.0, 1.5, 2.0), (2.2, 4.5, 1.0, 1.0)]):
        y0 = 1.5 - hgt
        fac.append(Facet([bx, y0, bz], [1, 0, 0], [0, 1, 0], s, hgt, amp=100, freq=7.0, tex_seed=seed + 10 + 3 * k))
        fac.append(Facet([bx, y0, bz], [1, 0, 0], [0, 0, 1], s, s, amp=100, freq=7.0, tex_seed=seed + 11 + 3 * k))
        fac.append(Facet([bx + s, y0, bz], [0, 0, 1], [0, 1, 0], s, hgt, amp=100, freq=7.0, tex_seed=seed + 12 + 3 * k))
        fac.append(Facet([bx, y0, bz], [0, 0, 1], [0, 1, 0], s, hgt, amp=100, freq=7.0, tex_seed=seed + 13 + 3 * k))
    return fac


def office_setup(width, height, num_views, seed=2, weak=0.0, arc_deg=60.0, rings=1):
    """facets, K and camera poses of the office scene: `rings` arcs of `num_views` cameras each, one above the other
    (0.3 m apart, closer than neighbours on an arc), views numbered ring by ring -- the nearest-camera source lists of a
    view then reach into the neighbouring rings, which a multi-GPU job hands to other ranks"""
    f = 0.85 * width
    K = np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1.0]])
    facets = _room_facets(seed, weak)
    target = np.array([0.0, 0.3, 5.5])
    poses = []
    radius = 5.5
    for g in range(rings):
        lift = 0.3 * (g - (rings - 1) / 2.0)
        for i in range(num_views):
            a = np.deg2rad(-arc_deg / 2 + arc_deg * i / max(1, num_views - 1))
            eye = target + radius * np.array([np.sin(a), -0.25 + 0.1 * np.sin(3 * a), -np.cos(a)]) + np.array([0.0, lift, 0.0])
            poses.append(look_at(eye, target))
    return facets, K, poses


def make_office_scene(width=1550, height=1030, num_views=26, num_src=10, seed=2, weak=0.0, arc_deg=60.0, with_color=False,
                      prerendered=None, rings=1):
    """C2 / C3 / C4 / C5 shaped: floor + two walls + three boxes, cameras on an arc (or `rings` arcs) looking at the room
    centre.  weak > 0 adds weak-texture blobs (amplitude <= 1 grey level) covering roughly that fraction of every surface."""
    facets, K, poses = office_setup(width, height, num_views, seed, weak, arc_deg, rings)
    return _finish(Scene(width, height), facets, K, poses, num_src, (2.0, 12.0), with_color, prerendered)


def depth_accuracy(depth, gt, rel=0.01):
    """fraction of ground-truth-valid pixels whose depth is within rel of the truth"""
    valid = gt > 0
    ok = np.abs(depth - gt) <= rel * gt
    return float((ok & valid).sum()) / max(1, int(valid.sum()))


def make_label_map(w, h, seed, zero_share=0.25):
    """synthetic segment-label map in the format of tools/run_SAM.py (uint8, 0 = no segment): blocky segments with ragged
    borders, label = 1 + cell index (mod 250), a share of the cells unlabelled, a few single-pixel segments"""
    rng = np.random.default_rng(seed)
    gy, gx = np.mgrid[0:h, 0:w]
    jx = (6 * np.sin(gy / 9.0 + seed)).astype(int)
    jy = (5 * np.cos(gx / 11.0 - seed)).astype(int)
    cell = ((gx + jx) // 23) + 13 * ((gy + jy) // 19)
    lut = (1 + np.arange(cell.max() + 1) % 250).astype(np.uint8)
    lut[rng.random(len(lut)) < zero_share] = 0
    lab = lut[cell]
    # a few single-pixel segments (a pixel whose four diagonal neighbours all differ: empty quadrant walk)
    ys, xs = rng.integers(8, h - 8, 40), rng.integers(8, w - 8, 40)
    lab[ys, xs] = 251
    return lab
