"""Post-processing / export methods of TensorProductSimulator that lie next to the solve path:
getMesh, sampleNodalField, getK, getDirichletVarsAndValues, getForceMask, getBCIndicatorField, constantStrainLoad,
solveWithImposedLoads, debugMulticolorElementVisit and the intermediate-fabrication-shape transfers.

Written against the PUBLIC simulator API only (plus the two accessors `_dirichletConditions()` / `_forceNodes()` both module
flavours provide), so the ctypes flavour (compat/pyVoxelFEM.py) and the pybind11 flavour (host/pyVoxelFEM.cc) bind the same code.
Host-side numpy: none of this is on the hot path.  All file:line citations refer to TensorProductSimulator.hh of the reference."""
import numpy as np


def _dims(tps):
    ne = np.asarray(tps.NbElementsPerDimension, dtype=np.int64)
    dmin, dmax = (np.asarray(a, dtype=np.float64) for a in tps.domain)
    return ne, ne + 1, dmin, dmax, (dmax - dmin) / ne


def node_positions(tps):
    """(numNodes, N) positions, node index row-major with the last axis fastest (nodePosition, :1700-1712)."""
    ne, nn, dmin, dmax, dx = _dims(tps)
    grids = np.meshgrid(*[dmin[d] + dx[d] * np.arange(nn[d]) for d in range(len(ne))], indexing="ij")
    return np.stack([g.ravel() for g in grids], axis=1)


def element_nodes(tps):
    """(numElements, 2^N) global node indices, local node n at offset bits (bit N-1-d of n <-> axis d) (:1532-1651)."""
    ne, nn, _, _, _ = _dims(tps)
    N = len(ne)
    first = np.ravel_multi_index(np.meshgrid(*[np.arange(n) for n in ne], indexing="ij"), tuple(nn)).ravel()
    strides = np.array([int(np.prod(nn[d + 1:])) for d in range(N)])
    offs = np.array([sum(((n >> (N - 1 - d)) & 1) * strides[d] for d in range(N)) for n in range(2 ** N)])
    return first[:, None] + offs[None, :]


def getMesh(tps):
    """(V, F): vertices (numNodes, 3) -- z = 0 in 2D, as MeshIO::IOVertex -- and elements in Gmsh node ordering: local vertex
    pairs (2p, 2p + 1), every odd pair swapped (getMesh, :747-777; binding VoxelFEM.cc:148-153)."""
    P = node_positions(tps)
    V = np.zeros((P.shape[0], 3)); V[:, :P.shape[1]] = P
    en = element_nodes(tps)
    order = []
    for pair in range(en.shape[1] // 2):
        order += [2 * pair, 2 * pair + 1] if pair % 2 == 0 else [2 * pair + 1, 2 * pair]
    return V, en[:, order]


def sampleNodalField(tps, u, p, reference_literal=False):
    """Sample the nodal field u (numNodes, k) at the rows of p (:1161-1179): closest-point projection onto the domain, element
    lookup with the boundary snapped back into the grid (getElementNDIndex, :1731-1752), multilinear interpolation.

    The reference computes the element's reference coordinates and then evaluates the interpolant at the clamped GLOBAL point
    (`Element::interpolate(nodeIndexGetter, u, q)`, :1175) -- correct only where the two coincide.  This function interpolates at
    the reference coordinates; reference_literal=True reproduces the reference's expression."""
    ne, nn, dmin, dmax, dx = _dims(tps)
    N = len(ne)
    u = np.asarray(u, dtype=np.float64)
    if u.ndim == 1: u = u[:, None]
    q = np.clip(np.atleast_2d(np.asarray(p, dtype=np.float64))[:, :N], dmin, dmax)
    fidx = (q - dmin) / dx
    e = np.where(np.abs(fidx - ne) < 1e-10, ne - 1, np.floor(fidx).astype(np.int64))
    if np.any(e >= ne): raise RuntimeError("Point out of bounds")
    ref = q if reference_literal else fidx - e
    out = np.zeros((q.shape[0], u.shape[1]))
    for n in range(2 ** N):
        off = np.array([(n >> (N - 1 - d)) & 1 for d in range(N)])
        w = np.prod(np.where(off == 1, ref, 1.0 - ref), axis=1)
        out += w[:, None] * u[np.ravel_multi_index(tuple((e + off).T), tuple(nn))]
    return out


def getDirichletVarsAndValues(tps):
    """(vars, values): variable N * node + c of every constrained component with its prescribed value (:1790-1835)."""
    nodes, masks, vals = tps._dirichletConditions()
    N = vals.shape[1] if vals.ndim == 2 and vals.shape[0] else len(np.asarray(tps.NbElementsPerDimension))
    v, x = [], []
    for ni, m, row in zip(nodes, masks, vals):
        for c in range(N):
            if (int(m) >> c) & 1:
                v.append(N * int(ni) + c); x.append(float(row[c]))
    return v, x


def getForceMask(tps):
    """(numNodes, N) bool: components carrying a non-zero nodal force (:685-695)."""
    N = len(np.asarray(tps.NbElementsPerDimension))
    out = np.zeros((tps.numNodes(), N), dtype=bool)
    nodes, f = tps._forceNodes()
    if len(nodes): out[nodes] = f != 0
    return out


def getBCIndicatorField(tps):
    """Per node: constrained-component bits + 2^N * forced-component bits (:697-712)."""
    N = len(np.asarray(tps.NbElementsPerDimension))
    out = np.zeros(tps.numNodes())
    nodes, masks, _ = tps._dirichletConditions()
    np.add.at(out, nodes, masks.astype(np.float64))
    fn, f = tps._forceNodes()
    if len(fn):
        bits = ((f != 0) * (1 << np.arange(N))).sum(axis=1)
        np.add.at(out, fn, (1 << N) * bits.astype(np.float64))
    return out


def debugMulticolorElementVisit(tps):
    """Visit rank of every element in visitElementsMulticolored's serial order (:1444-1457, 1482-1492): colours in
    HypercubeCornerVisitor order (axis 0 outermost), elements 2 k + colour offset row-major within a colour."""
    ne = np.asarray(tps.NbElementsPerDimension, dtype=np.int64)
    N = len(ne)
    out = np.zeros(int(np.prod(ne)))
    i = 0
    for col in range(2 ** N):
        off = np.array([(col >> (N - 1 - d)) & 1 for d in range(N)])
        if np.any(off >= ne): continue
        idx = np.meshgrid(*[np.arange(off[d], ne[d], 2) for d in range(N)], indexing="ij")
        flat = np.ravel_multi_index([g.ravel() for g in idx], tuple(ne))
        out[flat] = i + np.arange(flat.size); i += flat.size
    return out


def _strain_matrix(eps, N):
    e = np.asarray(eps, dtype=np.float64)
    if e.shape == (N, N): return 0.5 * (e + e.T)
    e = e.ravel()
    if e.size != N * (N + 1) // 2: raise RuntimeError("constantStrainLoad: eps must be an N x N matrix or the flattened symmetric matrix")
    m = np.zeros((N, N))
    # SymmetricMatrix flattening (MeshFEM SymmetricMatrix.hh:138-148): diagonal first, then yz, xz, xy in 3D / xy in 2D
    for i in range(N): m[i, i] = e[i]
    if N == 2: m[0, 1] = m[1, 0] = e[2]
    else:
        m[1, 2] = m[2, 1] = e[3]; m[0, 2] = m[2, 0] = e[4]; m[0, 1] = m[1, 0] = e[5]
    return m


def constantStrainLoad(tps, eps):
    """Global load of the constant unit strain eps (:1119-1150; Element::constantStrainLoad :107-113, constantStressLoad :85-100):
    F_n += rho_e * vol * int strain(phi_n e_i) : C : eps.  Multilinear elements reproduce the linear field u(x) = eps x exactly,
    so the element load is K0 u_lin(element) -- identical for every element up to the (stiffness-free) translation -- scaled by
    the element's DENSITY (not its interpolated modulus, as in the reference)."""
    ne, nn, dmin, dmax, dx = _dims(tps)
    N = len(ne)
    E = _strain_matrix(eps, N)
    K0 = np.asarray(tps.fullDensityElementStiffnessMatrix(), dtype=np.float64)
    loc = np.array([[((n >> (N - 1 - d)) & 1) * dx[d] for d in range(N)] for n in range(2 ** N)])
    le = (K0 @ (loc @ E.T).ravel()).reshape(2 ** N, N)          # load of one full-density element, per local node
    rho = np.asarray(tps.getDensities(), dtype=np.float64).reshape(tuple(ne))
    F = np.zeros(tuple(nn) + (N,))
    for n in range(2 ** N):
        sl = tuple(slice(((n >> (N - 1 - d)) & 1), ((n >> (N - 1 - d)) & 1) + ne[d]) for d in range(N))
        F[sl] += rho[..., None] * le[n]
    return F.reshape(-1, N)


def solveWithImposedLoads(tps):
    """solve(buildLoadVector()) (:1266)."""
    return tps.solve(tps.buildLoadVector())


def getK(tps):
    """The assembled stiffness matrix, upper triangle, as scipy.sparse.csc_matrix over variables N * node + c (getK :1506-1512,
    m_assembleStiffnessMatrix :834-865: K_e = E_e K0 accumulated over the elements, `di > dj` skipped).  Stands in for MeshFEM's
    SuiteSparseMatrix; meant for inspection and small direct solves, not for the solve path."""
    import scipy.sparse as sp
    ne, nn, _, _, _ = _dims(tps)
    N = len(ne)
    K0 = np.asarray(tps.fullDensityElementStiffnessMatrix(), dtype=np.float64)
    E = np.asarray(tps.getYoungModulusScaleFactor(), dtype=np.float64)
    en = element_nodes(tps)
    dofs = (N * en[:, :, None] + np.arange(N)[None, None, :]).reshape(en.shape[0], -1)      # (ne, N 2^N), local dof N n + c
    n = N * int(np.prod(nn))
    K = sp.csc_matrix((n, n))
    chunk = max(1, 2_000_000 // K0.size)
    for a in range(0, en.shape[0], chunk):
        d = dofs[a:a + chunk]
        rows = np.repeat(d, d.shape[1], axis=1).ravel(); cols = np.tile(d, (1, d.shape[1])).ravel()
        vals = (E[a:a + chunk, None] * K0.ravel()[None, :]).ravel()
        keep = rows <= cols
        K = K + sp.csc_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n, n))
    return K


def _intermediate_layers(tps, inter):
    ne, nei = np.asarray(tps.NbElementsPerDimension), np.asarray(inter.NbElementsPerDimension)
    ok = nei == ne; ok[1] = nei[1] <= ne[1]
    if not np.all(ok): raise RuntimeError("Intermediate shape is of unexpected size")
    return ne, nei


def transferVFieldToIntermediateFabricationShape(tps, intermediateTPS, u):
    """Rows of u on the node layers the intermediate shape has (:2010-2021)."""
    ne, nei = _intermediate_layers(tps, intermediateTPS)
    u = np.asarray(u, dtype=np.float64)
    return np.ascontiguousarray(u.reshape(tuple(ne + 1) + (u.shape[1],))[:, :nei[1] + 1]).reshape(-1, u.shape[1])


def accumElementScalarFieldFromIntermediateFabricationShape(tps, intermediateTPS, rho_in, rho_accum):
    """rho_accum[element of this simulator] += rho_in[element of the intermediate shape], IN PLACE (:2027-2034)."""
    ne, nei = _intermediate_layers(tps, intermediateTPS)
    acc = np.asarray(rho_accum)
    if acc.dtype != np.float64 or not acc.flags.c_contiguous or not acc.flags.writeable:
        raise RuntimeError("rho_accum must be a writable, contiguous float64 array (it is updated in place)")
    acc.reshape(tuple(ne))[:, :nei[1]] += np.asarray(rho_in, dtype=np.float64).reshape(tuple(nei))
