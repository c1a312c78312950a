"""Independent numpy/scipy restatement of the *mathematics* (not the reference's code)
used to pin the oracle: textbook B-matrix element stiffness, assembled sparse K,
multilinear prolongation matrix, explicit multicolour Gauss-Seidel.

Conventions: nodes flattened row-major (axis 0 slowest), local
node n = sum_d bit_d << (N-1-d), DOF = N*node + component.
"""
import itertools

import numpy as np
import scipy.sparse as sp


def elasticity_D(N, E, nu):
    """Engineering-strain Voigt matrix; 2D = plane stress (ElasticityTensor.hh:100-115)."""
    mu = E / (2 * (1 + nu))
    if N == 3:
        lam = nu * E / ((1 + nu) * (1 - 2 * nu))
        D = np.zeros((6, 6))
        D[:3, :3] = lam
        D[np.arange(3), np.arange(3)] = lam + 2 * mu
        D[3, 3] = D[4, 4] = D[5, 5] = mu
    else:
        lam = nu * E / (1 - nu * nu)
        D = np.array([[lam + 2 * mu, lam, 0], [lam, lam + 2 * mu, 0], [0, 0, mu]])
    return D


def k0_reference(N, h, E=1.0, nu=0.0, D=None):
    """K0 = int_e B^T D B with B in engineering strains, 3-point Gauss (over-integrated on purpose).  D: Voigt matrix in MeshFEM's
    flattening order (xx, yy[, zz, yz, xz], xy), entries C_ijkl; default isotropic (E, nu)."""
    D = elasticity_D(N, E, nu) if D is None else np.asarray(D, dtype=float)
    gp, gw = np.polynomial.legendre.leggauss(3)
    gp = 0.5 * (gp + 1)
    gw = 0.5 * gw
    npe = 2 ** N
    ke = N * npe
    K = np.zeros((ke, ke))
    bits = [[(n >> (N - 1 - d)) & 1 for d in range(N)] for n in range(npe)]
    for q in itertools.product(range(3), repeat=N):
        xi = [gp[i] for i in q]
        w = np.prod([gw[i] for i in q])
        grads = np.zeros((npe, N))
        for n in range(npe):
            for c in range(N):
                v = 1.0
                for d in range(N):
                    if d == c:
                        v *= (1.0 if bits[n][d] else -1.0) / h[d]
                    else:
                        v *= xi[d] if bits[n][d] else 1 - xi[d]
                grads[n, c] = v
        if N == 3:
            B = np.zeros((6, ke))
            for n in range(npe):
                gx, gy, gz = grads[n]
                B[0, 3 * n] = gx; B[1, 3 * n + 1] = gy; B[2, 3 * n + 2] = gz
                B[3, 3 * n + 1] = gz; B[3, 3 * n + 2] = gy   # gamma_yz
                B[4, 3 * n] = gz; B[4, 3 * n + 2] = gx       # gamma_xz
                B[5, 3 * n] = gy; B[5, 3 * n + 1] = gx       # gamma_xy
        else:
            B = np.zeros((3, ke))
            for n in range(npe):
                gx, gy = grads[n]
                B[0, 2 * n] = gx; B[1, 2 * n + 1] = gy
                B[2, 2 * n] = gy; B[2, 2 * n + 1] = gx
        K += w * B.T @ D @ B
    return K * np.prod(h)


def node_index_grid(nn):
    return np.arange(int(np.prod(nn))).reshape(nn)


def element_nodes(ne):
    """(numElems, 2^N) node indices, elements and nodes flattened row-major."""
    ne = np.asarray(ne)
    N = len(ne)
    nn = ne + 1
    idx = node_index_grid(nn)
    cols = []
    for n in range(2 ** N):
        sl = tuple(slice((n >> (N - 1 - d)) & 1, ((n >> (N - 1 - d)) & 1) + ne[d]) for d in range(N))
        cols.append(idx[sl].ravel())
    return np.stack(cols, axis=1)


def assemble_K(ne, K0, E):
    """Global stiffness in DOF order N*node + c (scipy CSR, full symmetric storage)."""
    ne = np.asarray(ne)
    N = len(ne)
    en = element_nodes(ne)
    npe = 2 ** N
    dofs = (en[:, :, None] * N + np.arange(N)[None, None, :]).reshape(len(en), npe * N)
    rows = np.repeat(dofs, npe * N, axis=1).ravel()
    cols = np.tile(dofs, (1, npe * N)).ravel()
    vals = (np.asarray(E).ravel()[:, None] * K0.ravel()[None, :]).ravel()
    ndof = int(np.prod(ne + 1)) * N
    return sp.csr_matrix((vals, (rows, cols)), shape=(ndof, ndof))


def field_to_dof(u):
    return np.asarray(u).reshape(-1)  # (numNodes, N) row-major == N*node + c


def dof_to_field(x, N):
    return np.asarray(x).reshape(-1, N)


def prolongation_1d(nc_elems):
    nf = 2 * nc_elems + 1
    nc = nc_elems + 1
    P = sp.lil_matrix((nf, nc))
    for i in range(nf):
        if i % 2 == 0:
            P[i, i // 2] = 1.0
        else:
            P[i, (i - 1) // 2] = 0.5
            P[i, (i + 1) // 2] = 0.5
    return P.tocsr()


def prolongation(ne_coarse, N_comp=None):
    """Fine-node x coarse-node multilinear interpolation, optionally kron'ed to DOFs."""
    P = None
    for d, n in enumerate(ne_coarse):
        Pd = prolongation_1d(int(n))
        P = Pd if P is None else sp.kron(P, Pd, format="csr")
    if N_comp:
        P = sp.kron(P, sp.identity(N_comp), format="csr")
    return P


def free_dof_mask(dmask, N):
    """dmask: per-node bitmask of constrained components -> boolean array over DOFs (True = free)."""
    bits = (dmask[:, None] >> np.arange(N)[None, :]) & 1
    return (bits == 0).reshape(-1)


def direct_solve(K, f, dmask, N):
    import scipy.sparse.linalg as spla
    free = free_dof_mask(dmask, N)
    x = np.zeros(K.shape[0])
    Kff = K[free][:, free].tocsc()
    x[free] = spla.spsolve(Kff, field_to_dof(f)[free])
    return dof_to_field(x, N)


def colored_gauss_seidel(K, u, b, nn, dmask, forward=True, nlimit=None):
    """One multicolour block-GS sweep (MultigridSolver.hh:347-378, 408-442) on the assembled matrix."""
    nn = np.asarray(nn)
    N = len(nn)
    u = np.array(u, dtype=float)
    Kc = K.tocsr()
    lim = np.array(nn if nlimit is None else nlimit)
    colors = range(2 ** N) if forward else range(2 ** N - 1, -1, -1)
    full = 2 ** N - 1
    for col in colors:
        off = [(col >> (N - 1 - d)) & 1 for d in range(N)]
        ranges = [range(off[d], lim[d], 2) for d in range(N)]
        # all nodes of a colour are updated from the same snapshot (Jacobi within a colour is
        # identical to GS within a colour because same-colour nodes are never coupled)
        for nd in itertools.product(*ranges):
            n = int(np.ravel_multi_index(nd, nn))
            if dmask[n] == full:
                continue
            dofs = np.arange(N * n, N * n + N)
            rhs = b[n] - Kc[dofs].dot(field_to_dof(u))
            M = Kc[dofs][:, dofs].toarray()
            if dmask[n] != 0:
                ud = np.zeros(N)
                order = range(N) if forward else range(N - 1, -1, -1)
                for i in order:
                    free = 0.0 if (dmask[n] >> i) & 1 else 1.0
                    ud[i] = (rhs[i] - M[i] @ ud) * (free / M[i, i])
                u[n] += ud
            else:
                u[n] += np.linalg.solve(M, rhs)
    return u
