"""Synthetic workloads for tests and bench.py — numpy only, every generator seeded (numpy PCG64).

Shapes of the inputs follow BASELINE.json's configs / SURVEY.md §8(d):
  maps        : u8 occupancy [X, Y, Z] (z fastest), three-slit walls (demo1 geometry) or Bernoulli + wall slabs
  trajectories: MINCO s=3 coefficient blocks (6N x 3, column-major) through random-walk waypoints 3 m apart
                (traj_parlength 3.0, plan_manager.cpp:153), T_i = inittime = 2.5 s
  robot meshes: procedurally generated closed triangle meshes (rounded cone ~4000 tris like RoundedCone.obj,
                L-shaped prism ~ Lthick.obj, box, icosphere). Nothing is read from the reference tree.
"""
import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# maps
def three_slit_map(X=64, Y=64, Z=64, noise=0.0, seed=0):
    """demo1's map semantics: three walls at x in {10-11, 24-26, 42-44} with slits (51x51x35 embedded in the grid)."""
    occ = np.zeros((X, Y, Z), dtype=np.uint8)
    walls = [(10, 12, (18, 30), (4, 22)), (24, 27, (28, 40), (8, 26)), (42, 45, (12, 24), (6, 24))]
    ymax, zmax = min(Y, 51), min(Z, 35)
    for x0, x1, (sy0, sy1), (sz0, sz1) in walls:
        if x1 > X:
            continue
        occ[x0:x1, 0:ymax, 0:zmax] = 1
        occ[x0:x1, sy0:sy1, sz0:sz1] = 0  # the slit
    if noise > 0:
        rng = np.random.default_rng(seed)
        occ |= (rng.random((X, Y, Z)) < noise).astype(np.uint8)
    return occ


def random_map(X, Y, Z, p=0.05, seed=0, slabs=2):
    """Bernoulli(p) occupancy plus a few axis-aligned wall slabs with a window each."""
    rng = np.random.default_rng(seed)
    occ = (rng.random((X, Y, Z), dtype=np.float32) < p).astype(np.uint8)
    for s in range(slabs):
        x = int((s + 1) * X / (slabs + 1))
        occ[x:x + 2, :, :] = 1
        wy, wz = int(rng.integers(Y // 8, Y - Y // 4)), int(rng.integers(Z // 8, Z - Z // 4))
        occ[x:x + 2, wy:wy + max(8, Y // 8), wz:wz + max(8, Z // 8)] = 0
    return occ


# ---------------------------------------------------------------------------------------------------------------
# trajectories
def random_walk_waypoints(N, lo, hi, seed=0, step=3.0, margin=8.0):
    """N+1 points, consecutive ones `step` apart, reflected at the box [lo+margin, hi-margin]."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, float) + margin
    hi = np.asarray(hi, float) - margin
    p = lo + (hi - lo) * rng.random(3)
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    pts = [p.copy()]
    for _ in range(N):
        d = d + 0.6 * rng.normal(size=3)
        d /= np.linalg.norm(d)
        q = p + step * d
        for a in range(3):
            if q[a] < lo[a] or q[a] > hi[a]:
                d[a] = -d[a]
        q = np.clip(p + step * d, lo, hi)
        pts.append(q.copy())
        p = q
    return np.array(pts)


def minco_s3(waypoints, T, head_va=None, tail_va=None):
    """Minimum-jerk quintic spline: dense numpy solve of the 6N x 6N MINCO system (independent of both the oracle's and
    the product's banded solvers). Returns coeffs as a flat array, 6N x 3 column-major (Eigen MatrixX3d)."""
    wp = np.asarray(waypoints, float)
    T = np.asarray(T, float)
    N = T.size
    A = np.zeros((6 * N, 6 * N))
    b = np.zeros((6 * N, 3))
    hv = np.zeros((2, 3)) if head_va is None else np.asarray(head_va, float)
    tv = np.zeros((2, 3)) if tail_va is None else np.asarray(tail_va, float)

    def row(t, der):
        r = np.zeros(6)
        for k in range(der, 6):
            c = 1.0
            for q in range(der):
                c *= (k - q)
            r[k] = c * t ** (k - der)
        return r
    A[0, 0:6] = row(0, 0); b[0] = wp[0]
    A[1, 0:6] = row(0, 1); b[1] = hv[0]
    A[2, 0:6] = row(0, 2); b[2] = hv[1]
    for i in range(N - 1):
        r = 6 * i
        A[r + 3, r:r + 6] = row(T[i], 3); A[r + 3, r + 6:r + 12] = -row(0, 3)
        A[r + 4, r:r + 6] = row(T[i], 4); A[r + 4, r + 6:r + 12] = -row(0, 4)
        A[r + 5, r:r + 6] = row(T[i], 0); b[r + 5] = wp[i + 1]
        A[r + 6, r:r + 6] = row(T[i], 0); A[r + 6, r + 6:r + 12] = -row(0, 0)
        A[r + 7, r:r + 6] = row(T[i], 1); A[r + 7, r + 6:r + 12] = -row(0, 1)
        A[r + 8, r:r + 6] = row(T[i], 2); A[r + 8, r + 6:r + 12] = -row(0, 2)
    e = 6 * N
    A[e - 3, e - 6:e] = row(T[-1], 0); b[e - 3] = wp[N]
    A[e - 2, e - 6:e] = row(T[-1], 1); b[e - 2] = tv[0]
    A[e - 1, e - 6:e] = row(T[-1], 2); b[e - 1] = tv[1]
    c = np.linalg.solve(A, b)  # (6N, 3)
    return np.ascontiguousarray(c.T).reshape(-1)  # column-major flat


def make_trajectory(N, lo, hi, seed=0, piece_time=2.5, jitter=0.0):
    wp = random_walk_waypoints(N, lo, hi, seed=seed)
    rng = np.random.default_rng(seed + 12345)
    T = piece_time * (1.0 + jitter * (rng.random(N) - 0.5))
    return T, minco_s3(wp, T), wp


def traj_eval(T, coeffs, t):
    """position at absolute time t (float or array) — plain helper for generators/tests."""
    T = np.asarray(T)
    N = T.size
    Cm = np.asarray(coeffs).reshape(3, 6 * N)
    t = np.atleast_1d(np.asarray(t, float))
    edges = np.concatenate([[0], np.cumsum(T)])
    out = np.zeros((t.size, 3))
    for q, tt in enumerate(t):
        i = min(np.searchsorted(edges, tt, side="right") - 1, N - 1)
        s = tt - edges[i]
        pw = s ** np.arange(6)
        out[q] = Cm[:, 6 * i:6 * i + 6] @ pw
    return out


def gather_obstacle_points(occ, bmin, res, waypoints, half):
    """parallel_points (plan_manager.cpp:232-254): occupied voxel centres in the union of boxes of half-extent `half`
    around the interior waypoints, deduplicated by voxel id, in ascending voxel-id order."""
    X, Y, Z = occ.shape
    bmin = np.asarray(bmin, float)
    ids = set()
    for w in np.asarray(waypoints)[1:-1]:
        lo = np.clip(np.floor((w - half - bmin) / res).astype(int), 0, [X - 1, Y - 1, Z - 1])
        hi = np.clip(np.floor((w + half - bmin) / res).astype(int), 0, [X - 1, Y - 1, Z - 1])
        sub = occ[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1]
        ix, iy, iz = np.nonzero(sub)
        for a, b_, c in zip(ix + lo[0], iy + lo[1], iz + lo[2]):
            ids.add((int(a) * Y + int(b_)) * Z + int(c))
    ids = np.array(sorted(ids), dtype=np.int64)
    if ids.size == 0:
        return np.zeros((0, 3))
    iz = ids % Z
    iy = (ids // Z) % Y
    ix = ids // (Y * Z)
    return np.stack([(ix + 0.5) * res + bmin[0], (iy + 0.5) * res + bmin[1], (iz + 0.5) * res + bmin[2]], axis=1)


# ---------------------------------------------------------------------------------------------------------------
# closed triangle meshes (outward orientation)
def box_mesh(hx=1.0, hy=0.5, hz=0.25):
    V = np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
    # vertex index = 4*ix + 2*iy + iz
    F = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6],
                  [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], np.int32)
    return V, F


def l_prism_mesh(a=3.0, b=2.0, t=0.8, h=0.6):
    """L-shaped thick plate (cf. Lthick.obj): polygon extruded along z, non-convex."""
    poly = np.array([[0, 0], [a, 0], [a, t], [t, t], [t, b], [0, b]], float) - np.array([a / 3, b / 3])
    n = len(poly)
    V = np.concatenate([np.c_[poly, -h / 2 * np.ones(n)], np.c_[poly, h / 2 * np.ones(n)]])
    tris2d = [[0, 1, 2], [0, 2, 3], [0, 3, 4], [0, 4, 5]]  # fan valid for this L (vertex 0 sees everything)
    F = []
    for tr in tris2d:
        F.append([tr[0], tr[2], tr[1]])              # bottom (normal -z)
        F.append([tr[0] + n, tr[1] + n, tr[2] + n])  # top (+z)
    for i in range(n):
        j = (i + 1) % n
        F.append([i, j, j + n])
        F.append([i, j + n, i + n])
    return V, np.array(F, np.int32)


def icosphere(radius=1.0, subdiv=2):
    t = (1.0 + 5 ** 0.5) / 2
    V = [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]]
    F = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
         [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]]
    V = [np.array(v, float) / np.linalg.norm(v) for v in V]
    for _ in range(subdiv):
        cache, F2 = {}, []

        def mid(i, j):
            key = (min(i, j), max(i, j))
            if key not in cache:
                m = V[i] + V[j]
                V.append(m / np.linalg.norm(m))
                cache[key] = len(V) - 1
            return cache[key]
        for a, b, c in F:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            F2 += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        F = F2
    return radius * np.array(V), np.array(F, np.int32)


def rounded_cone_mesh(r1=1.5, r2=0.6, h=4.5, n_theta=50, n_prof=40):
    """Surface of revolution of the RoundedCone profile (Shape.hpp:1003-1005 constants): sphere r1 at z=0, sphere r2
    at z=h, joined by the tangent cone. 2*n_theta*(n_prof-1) triangles (3900 for the defaults)."""
    b = (r1 - r2) / h
    a = np.sqrt(1 - b * b)
    # tangent points: on sphere 1 at angle where normal = (a, b) in (rho, z)
    n1 = n_prof // 2 - n_prof // 8
    n3 = n_prof // 4
    n2 = n_prof - n1 - n3
    phi_t = np.arctan2(b, a)  # elevation of the cone normal
    prof = []
    for s in np.linspace(-np.pi / 2, phi_t, n1, endpoint=False):      # bottom sphere cap, from south pole
        prof.append((r1 * np.cos(s), r1 * np.sin(s)))
    p1 = np.array([r1 * a, r1 * b]); p2 = np.array([r2 * a, h + r2 * b])
    for u in np.linspace(0, 1, n2, endpoint=False):                    # cone side
        prof.append(tuple(p1 + u * (p2 - p1)))
    for s in np.linspace(phi_t, np.pi / 2, n3):                        # top sphere cap, to north pole
        prof.append((r2 * np.cos(s), h + r2 * np.sin(s)))
    prof = np.array(prof)
    rho, z = prof[:, 0], prof[:, 1]
    m = len(prof)
    V = [[0, 0, z[0]]]
    th = np.linspace(0, 2 * np.pi, n_theta, endpoint=False)
    for i in range(1, m - 1):
        for tt in th:
            V.append([rho[i] * np.cos(tt), rho[i] * np.sin(tt), z[i]])
    V.append([0, 0, z[-1]])
    top = len(V) - 1

    def idx(i, j):
        return 1 + (i - 1) * n_theta + (j % n_theta)
    F = []
    for j in range(n_theta):
        F.append([0, idx(1, j + 1), idx(1, j)])
    for i in range(1, m - 2):
        for j in range(n_theta):
            F.append([idx(i, j), idx(i, j + 1), idx(i + 1, j + 1)])
            F.append([idx(i, j), idx(i + 1, j + 1), idx(i + 1, j)])
    for j in range(n_theta):
        F.append([idx(m - 2, j), idx(m - 2, j + 1), top])
    return np.array(V, float), np.array(F, np.int32)


def mesh_volume(V, F):
    """signed volume (positive for outward orientation)."""
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)


def rotation_from_poly_params(pp):
    """Rotate = Rz(yaw) Ry(pitch) Rx(roll), degrees (Shape.cpp:38-46); returns row-major 3x3 and trans."""
    r, p, y = np.deg2rad(pp[3]), np.deg2rad(pp[4]), np.deg2rad(pp[5])
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx, np.array(pp[:3], float)
