"""CPU pinning of the oracle's UpsampleFilter / VertexToCellFilter / LangelaarFilter restatements (oracle/vfo.cpp, namespace filt;
TopologyOptimizationFilter.hh:418-712) against independent numpy restatements, adjointness and finite differences."""
import numpy as np
import pytest

import oracle

RNG = np.random.default_rng(7)


def np_upsample(x, shape, f):
    """multilinear interpolation = linear interpolation along one axis after the other"""
    a = np.asarray(x, dtype=float).reshape(shape)
    for ax, s in enumerate(shape):
        fine = np.arange((s - 1) * f + 1) / f
        a = np.apply_along_axis(lambda v: np.interp(fine, np.arange(s), v), ax, a)
    return a.ravel()


@pytest.mark.parametrize("shape,f", [((3, 4), 2), ((2, 2), 3), ((5, 3), 4), ((3, 4, 2), 2), ((2, 3, 3), 3)])
def test_upsample(shape, f):
    x = RNG.normal(size=shape)
    y = oracle.upsample(x, shape, f)
    assert np.abs(y - np_upsample(x, shape, f)).max() < 1e-14
    fine = tuple((s - 1) * f + 1 for s in shape)
    assert np.array_equal(y.reshape(fine)[tuple(slice(None, None, f) for _ in shape)], x)       # coarse values are preserved (:407-409)
    g = RNG.normal(size=fine)
    assert abs(float(y @ g.ravel()) - float(x.ravel() @ oracle.upsample_backprop(g, shape, f))) < 1e-12 * np.abs(y).sum()   # backprop = transpose


@pytest.mark.parametrize("shape", [(3, 4), (2, 2), (4, 3, 5), (2, 2, 2)])
def test_vertex_to_cell(shape):
    x = RNG.normal(size=shape)
    N = len(shape)
    ref = np.zeros(tuple(s - 1 for s in shape))
    for b in range(2 ** N):
        sl = tuple(slice((b >> d) & 1, (b >> d) & 1 + s - 1 if False else ((b >> d) & 1) + s - 1) for d, s in enumerate(shape))
        ref += x[sl]
    ref *= 2.0 ** -N
    y = oracle.vertex_to_cell(x, shape)
    assert np.abs(y - ref.ravel()).max() < 1e-15
    g = RNG.normal(size=ref.shape)
    assert abs(float(y @ g.ravel()) - float(x.ravel() @ oracle.vertex_to_cell_backprop(g, shape))) < 1e-13 * np.abs(y).sum()


P, Q, EPS = 40.0, 40.0 - 1.58, 1e-4


def smin(a, b): return 0.5 * (a + b - np.sqrt((a - b) ** 2 + EPS) + np.sqrt(EPS))


def np_langelaar_2d(x):
    """Langelaar's overhang filter in 2D (layers along axis 1, support = the three voxels below)"""
    nx, ny = x.shape
    out = np.zeros_like(x); sm = np.zeros_like(x)
    out[:, 0] = x[:, 0]
    for y in range(1, ny):
        for i in range(nx):
            sup = [out[i, y - 1]] + [out[j, y - 1] for j in (i - 1, i + 1) if 0 <= j < nx]
            sm[i, y] = sum(v ** P for v in sup) ** (1 / Q)
            out[i, y] = smin(x[i, y], sm[i, y])
    return out, sm


def test_langelaar_2d_matches_formulas_and_gradient():
    shape = (7, 6)
    x = RNG.uniform(0.05, 1.0, shape)
    y, sm = oracle.langelaar(x, shape)
    yr, smr = np_langelaar_2d(x)
    assert np.abs(y - yr.ravel()).max() < 1e-14 and np.abs(sm - smr.ravel()).max() < 1e-14
    # a printable structure passes unchanged up to the smooth-min offset; an unsupported voxel is removed
    col = np.zeros(shape); col[3, :] = 1.0
    yc, _ = oracle.langelaar(col, shape)
    assert yc.reshape(shape)[3, 0] == 1.0 and yc.reshape(shape)[3, -1] > 0.9
    fl = np.zeros(shape); fl[2, 4] = 1.0
    assert oracle.langelaar(fl, shape)[0].reshape(shape)[2, 4] < 0.05
    # backprop is the exact transpose Jacobian in 2D: finite differences of <w, apply(x)>
    w = RNG.normal(size=shape)
    gb = oracle.langelaar_backprop(w, x, y, sm, shape)
    fd = np.zeros(x.size); h = 1e-6
    for k in range(x.size):
        xp, xm = x.ravel().copy(), x.ravel().copy(); xp[k] += h; xm[k] -= h
        fd[k] = (float(w.ravel() @ oracle.langelaar(xp, shape)[0]) - float(w.ravel() @ oracle.langelaar(xm, shape)[0])) / (2 * h)
    assert np.abs(gb - fd).max() < 1e-6 * max(1.0, np.abs(fd).max())


def np_support_3d(c, sz):
    """NDVector::visitSupportingRegion as written (NDVector.hh:211-229): the loop runs over the first N - 1 axes"""
    x, y, z = c
    cand = [(x, y - 1, z), (x - 1, y - 1, z), (x + 1, y - 1, z), (x, y - 2, z), (x, y, z)]
    return [q for q in cand if all(0 <= q[d] < sz[d] for d in range(3))]


def test_langelaar_3d_literal_restatement():
    """3D: the reference's support holds the voxel two layers below and the voxel itself (previous content of the output array)."""
    sz = (3, 4, 3)
    x = RNG.uniform(0.05, 1.0, sz); prev = RNG.uniform(0.0, 1.0, sz)
    y, sm = oracle.langelaar(x, sz, out_prev=prev)
    out = prev.copy(); smr = np.zeros(sz)
    out[:, 0, :] = x[:, 0, :]
    for l in range(1, sz[1]):
        for i in range(sz[0]):
            for j in range(sz[2]):
                smr[i, l, j] = sum(out[q] ** P for q in np_support_3d((i, l, j), sz)) ** (1 / Q)
                out[i, l, j] = smin(x[i, l, j], smr[i, l, j])
    assert np.abs(y - out.ravel()).max() < 1e-14 and np.abs(sm - smr.ravel()).max() < 1e-14
    # multipliers, literally (computeLagrangeMultipliers, :643-661)
    w = RNG.normal(size=sz)
    lam = np.zeros(sz)
    def D(i, k):
        S = sum(out[q] ** P for q in np_support_3d(i, sz))
        return 0.5 * (1 + (x[i] - smr[i]) * ((x[i] - smr[i]) ** 2 + EPS) ** -0.5) * (P * out[k] ** (P - 1) / Q * S ** (1 / Q - 1))
    for l in range(sz[1] - 1, -1, -1):
        lam[:, l, :] = w[:, l, :]
        if l < sz[1] - 1:
            for i in range(sz[0]):
                for j in range(sz[2]):
                    v = (i, l + 1, j)
                    for k in np_support_3d(v, sz):
                        lam[k] += lam[v] * D(v, k)
    ref = lam.copy()
    for l in range(1, sz[1]):
        ref[:, l, :] *= 0.5 * (1 - (x[:, l, :] - smr[:, l, :]) * ((x[:, l, :] - smr[:, l, :]) ** 2 + EPS) ** -0.5)
    gb = oracle.langelaar_backprop(w, x, y, sm, sz)
    assert np.abs(gb - ref.ravel()).max() < 1e-12 * max(1.0, np.abs(ref).max())
