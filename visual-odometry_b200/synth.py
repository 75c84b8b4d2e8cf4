"""Deterministic synthetic inputs for tests and bench.py (data generation only — no hot-path math).

The reference's generators (src/utils.cpp:8-34) seed mt19937 from std::random_device, so they
are not reproducible; these keep their DISTRIBUTIONS and fix the seeds.

NN maps use a counter-based integer hash evaluated with int64 arithmetic that is identical in
numpy (host) and torch (device), so a 1e8-row map can be generated on the GPU and any slice of it
regenerated bit-exactly on the host.
"""
import numpy as np

_M32 = 0xFFFFFFFF
_MUL = 0x45D9F3B


def _hash32(x):
    """x: int64 array/tensor with values in [0, 2^32).  Works for numpy and torch alike."""
    x = (((x >> 16) ^ x) * _MUL) & _M32
    x = (((x >> 16) ^ x) * _MUL) & _M32
    x = (x >> 16) ^ x
    return x


def _unit(h):
    """24 high bits of a 32-bit hash -> float32 in [-1, 1), exactly representable."""
    return (h >> 8)


def nn_map_rows_np(row0, row1, seed=1234, dim=10):
    """Rows [row0,row1) of the synthetic map as an (n, dim+1) float32 array; col 0 = float(id)
    (vo_complete.cpp:22), cols 1.. = U(-1,1) appearance (the range of world.dat)."""
    n = row1 - row0
    idx = (np.arange(row0, row1, dtype=np.int64)[:, None] * dim + np.arange(dim, dtype=np.int64)[None, :])
    h = _hash32((idx + seed * 2654435761) & _M32)
    out = np.empty((n, dim + 1), dtype=np.float32)
    out[:, 0] = np.arange(row0, row1, dtype=np.float32)
    out[:, 1:] = (_unit(h).astype(np.float32) * np.float32(2.0 ** -23)) - np.float32(1.0)
    return out


def nn_map_rows_torch(row0, row1, device, seed=1234, dim=10, out=None):
    """Same rows, generated on `device` with torch int64 ops (bit-identical to the numpy version)."""
    import torch

    n = row1 - row0
    idx = (torch.arange(row0, row1, dtype=torch.int64, device=device)[:, None] * dim
           + torch.arange(dim, dtype=torch.int64, device=device)[None, :])
    h = _hash32((idx + seed * 2654435761) & _M32)
    if out is None:
        out = torch.empty((n, dim + 1), dtype=torch.float32, device=device)
    out[:, 0] = torch.arange(row0, row1, dtype=torch.float32, device=device)
    out[:, 1:] = _unit(h).to(torch.float32) * (2.0 ** -23) - 1.0
    return out


def nn_map_torch(n_rows, device, seed=1234, dim=10, chunk=4_000_000):
    import torch

    out = torch.empty((n_rows, dim + 1), dtype=torch.float32, device=device)
    for r0 in range(0, n_rows, chunk):
        r1 = min(n_rows, r0 + chunk)
        nn_map_rows_torch(r0, r1, device, seed, dim, out=out[r0:r1])
    return out


def nn_query_plan(n_queries, n_rows, seed=1234):
    """Which map row each query is planted on (-1: fresh random query) and its noise class.
    50 % exact copies, 25 % copies + U(-0.01,0.01) noise per dim, 25 % fresh U(-1,1)."""
    q = np.arange(n_queries, dtype=np.int64)
    h = _hash32((q * 7919 + seed * 40503 + 12345) & _M32)
    target = (h % max(n_rows, 1)).astype(np.int64)
    cls = (q % 4)  # 0,1: exact   2: noisy   3: fresh
    target = np.where(cls == 3, -1, target)
    return target, cls


def nn_queries_np(n_queries, n_rows, seed=1234, dim=10, map_rows_fn=None):
    """(Q, dim+1) float32 queries + the planted target row per query (-1 = none planted).
    `map_rows_fn(rows)` returns the appearance of the given map rows; defaults to the hash map."""
    target, cls = nn_query_plan(n_queries, n_rows, seed)
    out = np.empty((n_queries, dim + 1), dtype=np.float32)
    out[:, 0] = np.arange(n_queries, dtype=np.float32)
    # fresh queries
    idx = (np.arange(n_queries, dtype=np.int64)[:, None] * dim + np.arange(dim, dtype=np.int64)[None, :])
    hf = _hash32((idx + (seed + 77) * 2654435761) & _M32)
    fresh = (_unit(hf).astype(np.float32) * np.float32(2.0 ** -23)) - np.float32(1.0)
    out[:, 1:] = fresh
    planted = target >= 0
    if planted.any() and n_rows > 0:
        rows = target[planted]
        if map_rows_fn is None:
            ridx = rows[:, None] * dim + np.arange(dim, dtype=np.int64)[None, :]
            hm = _hash32((ridx + seed * 2654435761) & _M32)
            app = (_unit(hm).astype(np.float32) * np.float32(2.0 ** -23)) - np.float32(1.0)
        else:
            app = map_rows_fn(rows)
        noisy = (cls[planted] == 2)
        noise = (fresh[planted] * np.float32(0.01)).astype(np.float32)
        app = np.where(noisy[:, None], (app + noise).astype(np.float32), app)
        out[planted, 1:] = app
    return out, target


# ---- reference-distribution geometry (src/utils.cpp:8-34) ----------------------------------------
def generate_isometry3f(rng, scale=1.0):
    """axis ~ U(-1,1)^3 normalised, angle ~ U(-1,1) rad, t ~ U(-1,1)^3 (utils.cpp:11-19)."""
    a = rng.uniform(-1.0, 1.0, 3)
    a /= np.linalg.norm(a)
    ang = rng.uniform(-1.0, 1.0) * scale
    t = rng.uniform(-1.0, 1.0, 3) * scale
    Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * (Kx @ Kx)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T.astype(np.float32)


def generate_points3d(rng, n):
    """x,y ~ U(-10,10), z = U(-10,10)*0.1+1 (utils.cpp:26-30)."""
    p = rng.uniform(-10.0, 10.0, (n, 3)).astype(np.float32)
    p[:, 2] = p[:, 2] * np.float32(0.1) + np.float32(1.0)
    return p


def frustum_points3d(rng, n, K, cols=640, rows=480, z_lo=0.5, z_hi=2.0, margin=40.0):
    """Points that project inside the image of the identity camera: pixel ~ U(image - margin),
    depth ~ U(z_lo, z_hi), back-projected through K (SURVEY.md §8d 'frustum-dist')."""
    u = rng.uniform(margin, cols - 1 - margin, n)
    v = rng.uniform(margin, rows - 1 - margin, n)
    z = rng.uniform(z_lo, z_hi, n)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    p = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], axis=1)
    return p.astype(np.float32)


def default_K(f=180.0):
    return np.array([[f, 0, 320], [0, f, 240], [0, 0, 1]], dtype=np.float32)


def project_np(K, T, pts, rows=480, cols=640, z_near=0, z_far=10):
    """float64 pinhole projection used ONLY to synthesise measurements (not a parity oracle)."""
    pc = pts.astype(np.float64) @ T[:3, :3].astype(np.float64).T + T[:3, 3].astype(np.float64)
    ph = pc @ K.astype(np.float64).T
    with np.errstate(divide="ignore", invalid="ignore"):
        uv = ph[:, :2] / ph[:, 2:3]
    ok = (pc[:, 2] <= z_far) & (pc[:, 2] >= z_near) & (uv[:, 0] >= 0) & (uv[:, 0] <= cols - 1) \
        & (uv[:, 1] >= 0) & (uv[:, 1] <= rows - 1)
    return uv.astype(np.float32), ok


def picp_problem(n, seed=42, dist="frustum", f=180.0, pose_scale=0.1, outlier_frac=0.0,
                 shuffle=False):
    """A PICP problem shaped like picp_solver_test.cpp:45-78: measurements taken from a hidden
    GT pose, solver starts from identity.  Returns dict(world, image, pairs, K, T_gt, cam kwargs).
    dist='ref' uses the reference generator verbatim (few points are visible);
    dist='frustum' makes >=90 % of the n points valid correspondences."""
    rng = np.random.RandomState(seed)
    K = default_K(f)
    if dist == "ref":
        world = generate_points3d(rng, n)
        T_gt = generate_isometry3f(rng, 1.0)
    else:
        world = frustum_points3d(rng, n, K)
        T_gt = generate_isometry3f(rng, pose_scale)
    uv_gt, ok_gt = project_np(K, T_gt, world)
    _, ok_id = project_np(K, np.eye(4, dtype=np.float32), world)
    keep = np.nonzero(ok_gt & ok_id)[0].astype(np.int32)  # visible in both (picp_solver_test.cpp:8-26)
    image = uv_gt.copy()
    image[~ok_gt] = -1.0
    if outlier_frac > 0:
        nb = int(len(keep) * outlier_frac)
        bad = rng.choice(keep, nb, replace=False)
        image[bad] += rng.uniform(-300, 300, (nb, 2)).astype(np.float32)
    pairs = np.stack([keep, keep], axis=1).astype(np.int32)
    if shuffle:
        pairs = pairs[rng.permutation(len(pairs))]
    return dict(world=world, image=image, pairs=pairs, K=K, T_gt=T_gt,
                rows=480, cols=640, z_near=0, z_far=10)


def two_view_problem(n, seed=7, f=150.0, noise=0.0):
    """Two views of frustum points (essential_picp_test.cpp:45-82 shape): returns K, X (pose of
    camera 1 expressed in camera 2, i.e. world->cam2 with world = cam1), p1, p2, corr, gt points."""
    rng = np.random.RandomState(seed)
    K = default_K(f)
    pts = frustum_points3d(rng, n, K)
    X = generate_isometry3f(rng, 0.2)
    # a mostly lateral baseline of 0.4-0.6 keeps the two-ray intersection well conditioned
    X[:3, 3] = np.array([rng.choice([-1.0, 1.0]) * rng.uniform(0.4, 0.6), rng.uniform(-0.1, 0.1),
                         rng.uniform(-0.1, 0.1)], dtype=np.float32)
    p1, ok1 = project_np(K, np.eye(4, dtype=np.float32), pts)
    p2, ok2 = project_np(K, X, pts)
    if noise > 0:
        p2 = (p2 + rng.normal(0, noise, p2.shape)).astype(np.float32)
    keep = np.nonzero(ok1 & ok2)[0].astype(np.int32)
    corr = np.stack([keep, keep], axis=1).astype(np.int32)
    return dict(K=K, X=X, p1=p1, p2=p2, corr=corr, points=pts)
