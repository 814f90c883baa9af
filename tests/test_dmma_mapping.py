"""Lane-level model of ax.cuh kernel v5 (ax_cg_affine_mma_kernel / ax_cg_mma_kernel): the fragment layout of
mma.sync.aligned.m8n8k4.row.col.f64 and the index arithmetic of the kernel, restated in numpy, must reproduce
w = D^T G D u of one 8^3 element (core/hmholtz.f:191-217 with G = c_ab * w_i w_j w_k).  This is the check the mapping was
designed against before its first GPU run; it pins the layout decisions documented in DESIGN.md section 3:

  lane = 4 g + t owns the k-columns (i = g, j = t) and (i = g, j = t + 4)
  D u      : A = D(i = g, m = 2t + s),          B = p[k][pi(g)][2t + s]   (one 16-byte word per lane and plane)
  u D^T    : A = the lane's OWN nodes p[k][t + 4s][g],  B = D(pi(g), t + 4s)
  D^T wr   : A = D(m = 2t + s, i = g),          B = wr[k][pi(g)][2t + s]
  D^T ws   : A = the lane's OWN ws values,      B = D(t + 4s, pi(g))
with pi(2q + s) = q + 4s, the j that C-fragment column 2q + s stands for."""
import numpy as np


def dmma884(c0, c1, a, b, g, t):
    """One warp-wide DMMA: A[row = lane/4][col = lane%4] = a, B[row = lane%4][col = lane/4] = b,
    C[row = lane/4][col = 2 (lane%4) + {0, 1}] = (c0, c1)  (PTX ISA, m8n8k4 .f64 fragments)."""
    A, B, Cm = np.zeros((8, 4)), np.zeros((4, 8)), np.zeros((8, 8))
    A[g, t], B[t, g] = a, b
    Cm[g, 2 * t], Cm[g, 2 * t + 1] = c0, c1
    Dm = A @ B + Cm
    return Dm[g, 2 * t], Dm[g, 2 * t + 1]


def kernel_model(D, wgt, P, c):
    lane = np.arange(32)
    g, t = lane >> 2, lane & 3
    pg = (g >> 1) + 4 * (g & 1)
    flat = P.reshape(-1)
    dA = [D[g, 2 * t + s] for s in (0, 1)]
    dB = [D[pg, t + 4 * s] for s in (0, 1)]
    dAt = [D[2 * t + s, g] for s in (0, 1)]
    dBt = [D[t + 4 * s, pg] for s in (0, 1)]
    wij = [wgt[g] * wgt[t + 4 * s] for s in (0, 1)]
    own = [t * 8 + g, (t + 4) * 8 + g]
    frag = pg * 8 + 2 * t
    pc = np.array([[flat[k * 64 + own[s]] for k in range(8)] for s in (0, 1)])
    swr, sws, wc = np.zeros(512), np.zeros(512), np.zeros((2, 8, 32))
    z = np.zeros(32)
    for k in range(8):                                   # phase 1
        bx, by = flat[k * 64 + frag], flat[k * 64 + frag + 1]
        ur = dmma884(z, z, dA[0], bx, g, t)
        ur = dmma884(*ur, dA[1], by, g, t)
        us = dmma884(z, z, pc[0, k], dB[0], g, t)
        us = dmma884(*us, pc[1, k], dB[1], g, t)
        for s in (0, 1):
            ut = sum(D[k, m] * pc[s, m] for m in range(8))
            W = wij[s] * wgt[k]
            swr[k * 64 + own[s]] = (c[0] * ur[s] + c[1] * us[s] + c[2] * ut) * W
            sws[k * 64 + own[s]] = (c[1] * ur[s] + c[3] * us[s] + c[4] * ut) * W
            wt = (c[2] * ur[s] + c[4] * us[s] + c[5] * ut) * W
            for m in range(8):
                wc[s, m] += D[k, m] * wt
    out = np.zeros(512)
    for k in range(8):                                   # phase 2
        bx, by = swr[k * 64 + frag], swr[k * 64 + frag + 1]
        a = dmma884(wc[0, k], wc[1, k], dAt[0], bx, g, t)
        a = dmma884(*a, dAt[1], by, g, t)
        a = dmma884(*a, sws[k * 64 + own[0]], dBt[0], g, t)
        a = dmma884(*a, sws[k * 64 + own[1]], dBt[1], g, t)
        out[k * 64 + own[0]], out[k * 64 + own[1]] = a
    return out.reshape(8, 8, 8)


def test_warp_per_element_dmma_mapping_reproduces_the_operator():
    import oracle
    case = oracle.Case(1, 1, 1, nx=8)
    D, wgt = np.asarray(case.D).reshape(8, 8), np.asarray(case.w)
    rng = np.random.default_rng(5)
    for _ in range(3):
        P, c = rng.standard_normal((8, 8, 8)), rng.standard_normal(6)     # P[k][j][i]
        ur = np.einsum("im,kjm->kji", D, P)
        us = np.einsum("jm,kmi->kji", D, P)
        ut = np.einsum("km,mji->kji", D, P)
        W3 = wgt[:, None, None] * wgt[None, :, None] * wgt[None, None, :]
        wr = (c[0] * ur + c[1] * us + c[2] * ut) * W3
        ws = (c[1] * ur + c[3] * us + c[4] * ut) * W3
        wt = (c[2] * ur + c[4] * us + c[5] * ut) * W3
        ref = np.einsum("mi,kjm->kji", D, wr) + np.einsum("mj,kmi->kji", D, ws) + np.einsum("mk,mji->kji", D, wt)
        got = kernel_model(D, wgt, P, c)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


def test_every_node_has_exactly_one_owner_and_fragment_loads_cover_a_plane_once():
    lane = np.arange(32)
    g, t = lane >> 2, lane & 3
    pg = (g >> 1) + 4 * (g & 1)
    own = np.concatenate([t * 8 + g, (t + 4) * 8 + g])
    assert sorted(own) == list(range(64))                                   # 2 nodes per lane, 64 nodes per plane
    frag = np.concatenate([pg * 8 + 2 * t, pg * 8 + 2 * t + 1])
    assert sorted(frag) == list(range(64))                                  # the 32 LDS.128 read the plane exactly once
    assert np.all((pg * 8 + 2 * t) % 2 == 0)                                # 16-byte aligned
    assert sorted(set(pg)) == list(range(8))                                # pi is a permutation of the j's
