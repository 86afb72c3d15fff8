"""Seeded synthetic inputs shared by the CPU and GPU tests (numpy only)."""
import numpy as np


def make_offsets(num_levels=16, base_resolution=16, desired_resolution=2048, log2_hashmap_size=19, input_dim=3,
                 align_corners=False):
    """Level offsets exactly as GridEncoder.__init__ computes them (reference gridencoder/grid.py:96-122)."""
    pls = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offs, off = [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * pls ** i))
        n = min(2 ** log2_hashmap_size, (res if align_corners else res + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offs.append(off)
        off += n
    offs.append(off)
    return np.array(offs, np.int32), float(np.log2(pls))


def look_at_rays(side, radius=1.3, theta_deg=70.0, phi_deg=30.0, fov_deg=55.0):
    """side x side pinhole rays from an orbit camera looking at the origin (get_rays / circle_poses style)."""
    th, ph = np.deg2rad(theta_deg), np.deg2rad(phi_deg)
    centre = np.array([radius * np.sin(th) * np.sin(ph), radius * np.cos(th), radius * np.sin(th) * np.cos(ph)])
    fwd = -centre / np.linalg.norm(centre)
    up0 = np.array([0.0, -1.0, 0.0])
    right = np.cross(fwd, up0); right /= np.linalg.norm(right)
    up = np.cross(right, fwd); up /= np.linalg.norm(up)
    R = np.stack([right, up, fwd], -1)
    focal = side / (2 * np.tan(np.deg2rad(fov_deg) / 2))
    j, i = np.meshgrid(np.arange(side) + 0.5, np.arange(side) + 0.5, indexing="ij")
    d = np.stack([(i - side / 2) / focal, (j - side / 2) / focal, np.ones_like(i)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays_d = (d @ R.T).astype(np.float32)
    rays_o = np.broadcast_to(centre.astype(np.float32), rays_d.shape).copy()
    return rays_o, rays_d


def _morton_invert(m):
    def compact(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xc30c30c3
        x = (x | (x >> 4)) & 0x0f00f00f
        x = (x | (x >> 8)) & 0xff0000ff
        x = (x | (x >> 16)) & 0x0000ffff
        return x
    m = m.astype(np.uint32)
    return compact(m), compact(m >> 1), compact(m >> 2)


def blob_density_grid(cascade=1, H=128, bound=1.0, seed=0, speckle=0.002):
    """Morton-ordered density grid [cascade, H^3]: the reference's initial Gaussian blob
    (exp(5 exp(-|x|^2/0.08)), network_grid.py:66-84) plus a few random occupied speckles."""
    rng = np.random.default_rng(seed)
    m = np.arange(H ** 3)
    x, y, z = _morton_invert(m)
    grid = np.zeros((cascade, H ** 3), np.float32)
    for c in range(cascade):
        b = min(2.0 ** c, bound)
        pts = (np.stack([x, y, z], -1).astype(np.float64) + 0.5) / H * 2 - 1
        pts *= b
        d2 = (pts ** 2).sum(-1)
        grid[c] = np.exp(5 * np.exp(-d2 / 0.08)).astype(np.float32)
        sp = rng.random(H ** 3) < speckle
        grid[c][sp] += 50.0
    return grid


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# ---------------------------------------------------------------------------------------------------
# Golden cases: seeded inputs shared by oracle/make_golden.py (runs the reference's CUDA extensions on a
# GPU box and stores their OUTPUTS in tests/golden/ref_golden.npz), the CPU oracle tests and the GPU tests.
# ---------------------------------------------------------------------------------------------------
def golden_grid_case(gridtype, dtype, seed=11, B=384):
    """cfg3-shaped (tiled, 2^16) or cfg2-shaped-but-smaller (hash, 2^14) encoder problem."""
    rng = np.random.default_rng(seed)
    log2 = 16 if gridtype == 1 else 14
    offs, S = make_offsets(log2_hashmap_size=log2)
    emb = rng.uniform(-1, 1, (offs[-1], 2)).astype(np.float32)
    x = rng.uniform(0, 1, (B, 3)).astype(np.float32)
    x[0] = 0.0
    x[1] = 1.0
    x[2] = [1.25, 0.5, 0.5]
    g = rng.standard_normal((B, 32)).astype(np.float32)
    if dtype == "f16":
        emb = emb.astype(np.float16)
        g = g.astype(np.float16)
    return dict(x=x, emb=emb, offs=offs, S=np.float32(S), H=16, grad=g, gridtype=gridtype)


def golden_march_case(seed=12, side=12, max_steps=256, cascade=1, bound=1.0, dt_gamma=0.0):
    rays_o, rays_d = look_at_rays(side, radius=1.25 * bound, theta_deg=65, phi_deg=200)
    grid = blob_density_grid(cascade, 128, bound, seed)
    noises = np.random.default_rng(seed + 1).random(rays_o.shape[0]).astype(np.float32)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    return dict(rays_o=rays_o, rays_d=rays_d, grid=grid, thresh=10.0, aabb=aabb, noises=noises, bound=bound,
                cascade=cascade, max_steps=max_steps, dt_gamma=dt_gamma)


def pseudo_field(xyzs):
    """Deterministic smooth stand-in for the network: sigma, rgb as functions of position (fp32 numpy)."""
    x = xyzs.astype(np.float32)
    sigma = (np.float32(30.0) * np.exp(-(x * x).sum(-1) / np.float32(0.08))).astype(np.float32)
    rgb = (np.float32(0.5) + np.float32(0.5) * np.sin(np.float32(7.0) * x)).astype(np.float32)
    return sigma, rgb
