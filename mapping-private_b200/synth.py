"""Deterministic synthetic clouds for the BASELINE.json configurations (C1..C5).

Every coordinate is rounded to a multiple of 2**-16 m (~15 um) and stored as
float32, so coordinate differences and the fp32 squared distance of the
documented epsilon rule are exactly representable for |delta| <= 6 cm
(SURVEY.md section 8d).  Shapes follow the seven classes of the reference's
fixtures (color_chlac/demos/shape_data: cone, cube, cylinder, dice, plane,
sphere, torus).

PRNG: numpy PCG64 seeded with 0xC10D0000 + config id.
"""
from __future__ import annotations

import numpy as np

LATTICE = 65536.0
SEED_BASE = 0xC10D0000


def quantize(p: np.ndarray) -> np.ndarray:
    """Round to the 2**-16 m lattice and return float32 (exact for |p| < 128 m)."""
    return (np.rint(np.asarray(p, dtype=np.float64) * LATTICE) / LATTICE).astype(np.float32)


def _rng(config_id: int, extra: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([SEED_BASE + config_id, extra]))


# ----------------------------------------------------------------------------
# primitives: each returns (area, sampler(n, rng) -> (n,3) float64)
# ----------------------------------------------------------------------------
def rect(origin, eu, ev):
    origin, eu, ev = (np.asarray(a, dtype=np.float64) for a in (origin, eu, ev))
    area = float(np.linalg.norm(np.cross(eu, ev)))

    def sample(n, rng):
        u = rng.random(n)[:, None]
        v = rng.random(n)[:, None]
        return origin + u * eu + v * ev

    return area, sample


def box_faces(center, size):
    cx, cy, cz = center
    sx, sy, sz = size
    x0, y0, z0 = cx - sx / 2, cy - sy / 2, cz - sz / 2
    return [
        rect((x0, y0, z0), (sx, 0, 0), (0, sy, 0)),
        rect((x0, y0, z0 + sz), (sx, 0, 0), (0, sy, 0)),
        rect((x0, y0, z0), (sx, 0, 0), (0, 0, sz)),
        rect((x0, y0 + sy, z0), (sx, 0, 0), (0, 0, sz)),
        rect((x0, y0, z0), (0, sy, 0), (0, 0, sz)),
        rect((x0 + sx, y0, z0), (0, sy, 0), (0, 0, sz)),
    ]


def cylinder_side(center, radius, height):
    center = np.asarray(center, dtype=np.float64)
    area = 2 * np.pi * radius * height

    def sample(n, rng):
        t = rng.random(n) * 2 * np.pi
        h = (rng.random(n) - 0.5) * height
        return center + np.stack([radius * np.cos(t), radius * np.sin(t), h], axis=1)

    return area, sample


def sphere(center, radius):
    center = np.asarray(center, dtype=np.float64)
    area = 4 * np.pi * radius * radius

    def sample(n, rng):
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        return center + radius * v

    return area, sample


def cone_side(apex_center, radius, height):
    """Upright cone, base centre at apex_center - (0,0,height/2), apex on top."""
    c = np.asarray(apex_center, dtype=np.float64)
    slant = np.hypot(radius, height)
    area = np.pi * radius * slant

    def sample(n, rng):
        s = np.sqrt(rng.random(n))  # area-uniform along the slant
        t = rng.random(n) * 2 * np.pi
        rr = radius * s
        z = height / 2 - height * s
        return c + np.stack([rr * np.cos(t), rr * np.sin(t), z], axis=1)

    return area, sample


def torus(center, R, r):
    c = np.asarray(center, dtype=np.float64)
    area = 4 * np.pi * np.pi * R * r

    def sample(n, rng):
        out = np.empty((0, 3))
        while out.shape[0] < n:
            m = int((n - out.shape[0]) * 1.5) + 16
            u = rng.random(m) * 2 * np.pi
            v = rng.random(m) * 2 * np.pi
            keep = rng.random(m) * (R + r) <= R + r * np.cos(v)  # area element ~ (R + r cos v)
            u, v = u[keep], v[keep]
            p = np.stack([(R + r * np.cos(v)) * np.cos(u), (R + r * np.cos(v)) * np.sin(u), r * np.sin(v)], axis=1)
            out = np.concatenate([out, p])
        return c + out[:n]

    return area, sample


def sample_by_area(prims, n, rng):
    areas = np.array([a for a, _ in prims], dtype=np.float64)
    counts = rng.multinomial(n, areas / areas.sum())
    parts = [s(int(c), rng) for (_, s), c in zip(prims, counts) if c > 0]
    return np.concatenate(parts, axis=0)


# ----------------------------------------------------------------------------
# shape classes of the reference fixtures, unit scale ~ 10 cm
# ----------------------------------------------------------------------------
SHAPE_CLASSES = ("cone", "cube", "cylinder", "dice", "plane", "sphere", "torus")


def shape_prims(name: str, s: float = 1.0):
    if name == "cone":
        return [cone_side((0, 0, 0), 0.05 * s, 0.12 * s)]
    if name == "cube":
        return box_faces((0, 0, 0), (0.10 * s, 0.06 * s, 0.15 * s))
    if name == "cylinder":
        return [cylinder_side((0, 0, 0), 0.04 * s, 0.12 * s)]
    if name == "dice":
        return box_faces((0, 0, 0), (0.08 * s, 0.08 * s, 0.08 * s))
    if name == "plane":
        return [rect((-0.06 * s, -0.06 * s, 0), (0.12 * s, 0, 0), (0, 0.12 * s, 0))]
    if name == "sphere":
        return [sphere((0, 0, 0), 0.05 * s)]
    if name == "torus":
        return [torus((0, 0, 0), 0.045 * s, 0.01 * s)]
    raise ValueError(name)


def random_rotation(rng) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


# ----------------------------------------------------------------------------
# C1: tabletop, 100 k points
# ----------------------------------------------------------------------------
def tabletop(n: int = 100_000, noise_sigma: float = 0.0, seed_extra: int = 0) -> np.ndarray:
    """Table plane 1.0 x 0.6 m at z = 0.75 (60 %) plus five objects (8 % each)."""
    rng = _rng(1, seed_extra)
    n_obj = int(n * 0.08)
    n_table = n - 5 * n_obj
    parts = [sample_by_area([rect((-0.5, -0.3, 0.75), (1.0, 0, 0), (0, 0.6, 0))], n_table, rng)]
    objs = [
        ([sphere((-0.32, 0.05, 0.80), 0.05)]),
        ([cylinder_side((-0.15, -0.08, 0.81), 0.04, 0.12)]),
        (box_faces((0.03, 0.06, 0.825), (0.10, 0.06, 0.15))),
        ([cone_side((0.20, -0.07, 0.81), 0.05, 0.12)]),
        ([torus((0.36, 0.08, 0.76), 0.045, 0.01)]),
    ]
    for prims in objs:
        parts.append(sample_by_area(prims, n_obj, rng))
    p = np.concatenate(parts, axis=0)
    if noise_sigma > 0:
        p = p + rng.normal(scale=noise_sigma, size=p.shape)
    p = p[rng.permutation(p.shape[0])]
    return quantize(p)


# ----------------------------------------------------------------------------
# C2 / C4: room scene
# ----------------------------------------------------------------------------
def room_prims():
    prims = box_faces((0.0, 0.0, 1.25), (5.0, 4.0, 2.5))  # the shell (inner faces of the room)
    boxes = [
        ((-1.6, -1.2, 0.40), (0.8, 0.5, 0.8)),
        ((1.4, -1.3, 0.45), (0.6, 0.6, 0.9)),
        ((-1.2, 1.3, 0.35), (1.0, 0.4, 0.7)),
        ((1.7, 1.1, 0.50), (0.5, 0.7, 1.0)),
        ((0.2, -0.4, 0.375), (1.2, 0.7, 0.75)),
        ((-0.3, 1.5, 0.90), (0.4, 0.3, 1.8)),
    ]
    for c, s in boxes:
        prims += box_faces(c, s)
    cyls = [((-2.0, 0.3, 0.5), 0.15, 1.0), ((2.0, -0.2, 0.6), 0.12, 1.2), ((0.6, 1.2, 0.4), 0.20, 0.8), ((-0.7, -1.5, 0.55), 0.10, 1.1)]
    for c, r, h in cyls:
        prims.append(cylinder_side(c, r, h))
    return prims


def room(n: int = 20_000_000, config_id: int = 4) -> np.ndarray:
    """C4: uniform-area sampling of the room scene (~2e5 pts/m^2 at 20 M points)."""
    rng = _rng(config_id)
    chunks = []
    prims = room_prims()
    step = 4_000_000
    for s in range(0, n, step):
        chunks.append(quantize(sample_by_area(prims, min(step, n - s), rng)))
    p = np.concatenate(chunks, axis=0)
    return p[rng.permutation(p.shape[0])]


def scan(n: int = 1_000_000) -> np.ndarray:
    """C2: the room scene seen from a sensor at (0, 0, 1.2): area samples are kept with
    probability ~ 1/d^2 (capped), so density falls with range like an angular scan."""
    rng = _rng(2)
    prims = room_prims()
    sensor = np.array([0.0, 0.0, 1.2])
    out = []
    have = 0
    while have < n:
        cand = sample_by_area(prims, 2 * n, rng)
        d2 = np.sum((cand - sensor) ** 2, axis=1)
        keep = rng.random(cand.shape[0]) < np.minimum(1.0, 0.25 / d2)
        cand = cand[keep]
        out.append(cand)
        have += cand.shape[0]
    p = np.concatenate(out, axis=0)[:n]
    p = p[rng.permutation(n)]
    return quantize(p)


# ----------------------------------------------------------------------------
# C3: batch of segmented object clusters
# ----------------------------------------------------------------------------
def clusters(n_clusters: int = 512, min_pts: int = 1300, max_pts: int = 15000, seed_extra: int = 0):
    """Returns (xyz float32 (N,3), offsets int32 (n_clusters+1)).  Shape class = i mod 7,
    scale U[0.7,1.5], random SO(3) pose, 0.5 mm Gaussian noise on odd clusters."""
    rng = _rng(3, seed_extra)
    parts = []
    offsets = [0]
    for i in range(n_clusters):
        name = SHAPE_CLASSES[i % 7]
        s = rng.uniform(0.7, 1.5)
        npts = int(rng.integers(min_pts, max_pts + 1))
        p = sample_by_area(shape_prims(name, s), npts, rng)
        if i % 2 == 1:
            p = p + rng.normal(scale=0.0005, size=p.shape)
        R = random_rotation(rng)
        t = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.3, 0.3), rng.uniform(0.6, 1.2)])
        p = p @ R.T + t
        parts.append(quantize(p))
        offsets.append(offsets[-1] + npts)
    return np.concatenate(parts, axis=0), np.asarray(offsets, dtype=np.int32)


# ----------------------------------------------------------------------------
# C5: density sweep
# ----------------------------------------------------------------------------
def density_patches(n: int = 5_000_000, k_mean: float = 250.0, r: float = 0.02) -> np.ndarray:
    """Planar + curved patches whose area gives mean k = rho*pi*r^2 at radius r."""
    rng = _rng(5, int(k_mean))
    rho = k_mean / (np.pi * r * r)
    area = n / rho
    # half of the area planar (one tilted square), half on cylinders of radius 0.5 m
    side = np.sqrt(area / 2)
    c, s = np.cos(0.3), np.sin(0.3)
    prims = [rect((0, 0, 0), (side * c, 0, side * s), (0, side, 0))]
    cyl_area = area / 2
    n_cyl = max(1, int(np.ceil(cyl_area / (2 * np.pi * 0.5 * 2.0))))
    h = cyl_area / n_cyl / (2 * np.pi * 0.5)
    for i in range(n_cyl):
        prims.append(cylinder_side((-1.5 - 1.5 * (i % 8), 1.5 * (i // 8), h / 2), 0.5, h))
    p = sample_by_area(prims, n, rng)
    p = p[rng.permutation(n)]
    return quantize(p)


def analytic_shape(name: str, n: int, seed_extra: int = 0, **kw) -> np.ndarray:
    """Small analytic known-answer clouds: 'plane', 'sphere' (R), 'cylinder' (R, h)."""
    rng = _rng(9, seed_extra)
    if name == "plane":
        prims = [rect((0.5, 0.5, 1.0), (kw.get("side", 0.3), 0, 0), (0, kw.get("side", 0.3), 0))]
    elif name == "sphere":
        prims = [sphere((0.5, 0.5, 1.0), kw.get("R", 0.05))]
    elif name == "cylinder":
        prims = [cylinder_side((0.5, 0.5, 1.0), kw.get("R", 0.04), kw.get("h", 0.3))]
    else:
        raise ValueError(name)
    return quantize(sample_by_area(prims, n, rng))
