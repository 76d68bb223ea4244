"""CPU pins for the front-end oracle (oracle/oracle_frontend.hpp): its byte-packed kernelConv (the reference's formulation,
sw_manager.hpp:821-846 over PCSmap_manager.h:46-78 and Shape.hpp:232-256) must equal the plain definition — attitude (i,j) fits
at voxel v iff no kernel voxel (a,b,c) coincides with an occupied map voxel v + (a,b,c) - side, cells outside the map free —
evaluated independently with numpy; the BFS visiting order is checked against a direct restatement."""
import collections
import numpy as np
import oracle_lib as O
import workloads as W


def brute_masks(K, occ, ind, ks):
    side = (ks - 1) // 2
    pad = np.zeros(tuple(d + 2 * side for d in occ.shape), dtype=bool)
    pad[side:-side, side:-side, side:-side] = occ != 0
    out = np.zeros((len(ind), 4), dtype=np.uint32)
    for q, (x, y, z) in enumerate(ind):
        win = pad[x:x + ks, y:y + ks, z:z + ks]
        hit = (K.astype(bool) & win[None]).reshape(K.shape[0], -1).any(axis=1)
        for att in np.nonzero(~hit)[0]:
            out[q, att >> 5] |= np.uint32(1 << (att & 31))
    return out


def test_byte_conv_equals_definition():
    occ = W.three_slit_map(40, 36, 28, noise=0.004, seed=5)
    rng = np.random.default_rng(1)
    ind = np.stack([rng.integers(0, 40, 600), rng.integers(0, 36, 600), rng.integers(0, 28, 600)], 1)
    ind[:8] = [[0, 0, 0], [39, 35, 27], [0, 35, 0], [39, 0, 27], [20, 18, 0], [20, 18, 27], [0, 18, 14], [39, 18, 14]]   # corners, faces
    for name, ks in [("CappedCone", 13), ("TwistBox", 13), ("Torus", 9), ("Ball", 5)]:
        fe = O.FrontEnd(O.Shape.named(name), occ, ks=ks)
        K = fe.kernels()
        assert K.shape == (121, ks, ks, ks) and 0 < K[60].sum() < ks ** 3
        got = fe.feasibility(ind)
        assert np.array_equal(got, brute_masks(K, occ, ind, ks)), name
        assert (got != 0).any() and not (got[:, 0] == 0xffffffff).all()


def test_kernels_follow_attitude_and_margin():
    occ = np.zeros((8, 8, 8), dtype=np.uint8)
    sh = O.Shape.named("Torus")                      # a ring in the x-z plane: rolling it changes the footprint, the level kernel is symmetric
    fe = O.FrontEnd(sh, occ, ks=13)
    K = fe.kernels().astype(bool)
    lvl = K[60]
    assert np.array_equal(lvl, lvl[::-1, :, :]) and np.array_equal(lvl, lvl[:, :, ::-1])
    assert not np.array_equal(K[0], lvl)
    # the level kernel is "sdf <= res/2" on the voxel-centre lattice
    c = np.arange(13) - 6.0
    P = np.stack(np.meshgrid(c, c, c, indexing="ij"), -1).reshape(-1, 3)
    s, _ = sh.query(P, what=0)
    assert np.array_equal(lvl.reshape(-1), s <= 0.5)
    fe2 = O.FrontEnd(sh, occ, ks=13, front_end_safeh=1.2)
    assert fe2.kernels()[60].sum() > lvl.sum()       # a larger safety margin inflates the kernel


def bfs_order(xk, yk, sx, sy):
    zi, zj = (xk - 1) // 2, (yk - 1) // 2
    order, seen, q = [(zi, zj)], {(sx, sy)}, collections.deque([(sx, sy)])
    while q:
        x, y = q.popleft()
        if (x, y) != (zi, zj):
            order.append((x, y))
        for dx, dy in ((0, 1), (0, -1), (1, 0), (-1, 0)):
            n = (x + dx, y + dy)
            if 0 <= n[0] < xk and 0 <= n[1] < yk and n not in seen:
                seen.add(n); q.append(n)
    return order


def test_check_returns_first_fit_in_visiting_order():
    occ = W.three_slit_map(40, 36, 28, noise=0.002, seed=6)
    fe = O.FrontEnd(O.Shape.named("CappedCone"), occ, ks=13)
    rng = np.random.default_rng(2)
    n = 300
    ind = np.stack([rng.integers(0, 40, n), rng.integers(0, 36, n), rng.integers(0, 28, n)], 1)
    father = np.stack([rng.integers(-5, 6, n) * 9.0, rng.integers(-5, 6, n) * 9.0], 1)
    child, ok = fe.check(ind, father)
    masks = fe.feasibility(ind)
    some_moved = False
    for q in range(n):
        fi, fj = int((father[q, 0] + 45) / 9), int((father[q, 1] + 45) / 9)
        exp = None
        for (i, j) in bfs_order(11, 11, fi, fj):
            att = i * 11 + j
            if (int(masks[q, att >> 5]) >> (att & 31)) & 1:
                exp = (father[q, 0] + (i - fi) * 9.0, father[q, 1] + (j - fj) * 9.0)
                break
        assert ok[q] == (exp is not None)
        if exp is not None:
            assert tuple(child[q]) == exp
            some_moved |= tuple(child[q]) not in ((0.0, 0.0), tuple(father[q]))
    assert ok.any() and not ok.all() and some_moved
