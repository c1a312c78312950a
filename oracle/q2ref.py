"""CPU restatement (numpy) of the reference's GENERIC tensor-product element path for degree-2 (Q2) elements -- test
infrastructure only: nothing under voxelfem_b200/ may import it.

What it follows (all in /root/reference):
  * 1D Lagrange basis on equidistant nodes of [0, 1] (LagrangePolynomial.hh:7-57);
  * tensor-product shape functions, local node index row-major over (deg + 1)^N with the last axis fastest
    (TPSStencils.hh:139 ElementNodeIndexer = NDArrayIndexer<N, (Degrees + 1)...>);
  * strains of the vector-valued shape function  phi_(N m + c) = N_m e_c :  eps_(c i) = 1/2 d_i N_m, eps_(c c) = d_c N_m
    (TensorProductPolynomialInterpolant.hh:204-231);
  * K0 = vol * int eps_a : C : eps_b with a Gauss rule of degree 2 deg per axis -- 3 points for Q2, exact
    (TensorProductSimulator.hh:50, 67-80; TensorProductQuadrature.hh:129-160);
  * node grid (deg * ne + 1)^N, global node of local node m of element e = deg * e + m per axis, flat indices row-major with
    the last axis fastest (NDVector.hh:249-275; TensorProductSimulator.hh:1532-1651);
  * applyK as the element scatter of TPSStencils.hh:163-185:  f[nodes(e)] += E_e K0 u[nodes(e)].
PARITY UNPINNED by the reference: its bindings never instantiate degree 2 (python_bindings/VoxelFEM.cc:303-308) and it ships no
outputs; the restatement is pinned by invariants instead (tests/test_oracle_q2.py: symmetry, rigid-body null space, exact
energies of linear and quadratic displacement fields, agreement with the Q1 oracle on what both can represent)."""
import itertools

import numpy as np


def lagrange(deg, x):
    """Values and derivatives (deg + 1,) of the 1D Lagrange basis on the nodes j / deg at x."""
    nodes = np.arange(deg + 1) / deg
    val, der = np.ones(deg + 1), np.zeros(deg + 1)
    for i in range(deg + 1):
        others = [j for j in range(deg + 1) if j != i]
        w = 1.0 / np.prod([nodes[i] - nodes[j] for j in others])
        val[i] = w * np.prod([x - nodes[j] for j in others])
        der[i] = w * sum(np.prod([x - nodes[j] for j in others if j != k]) for k in others)
    return val, der


def gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def sym_idx(N, i, j):
    if i == j: return i
    return 2 if N == 2 else 6 - i - j


def isotropic_tensor(N, young, nu):
    """Flattened elasticity tensor in the repo's convention (vf_sim::setIsotropic): plane stress in 2D."""
    mu = young / (2.0 + 2.0 * nu)
    lam = nu * young / ((1.0 + nu) * (1.0 - 2.0 * nu)) if N == 3 else nu * young / (1.0 - nu * nu)
    fl = 6 if N == 3 else 3
    D = np.zeros((fl, fl))
    D[:N, :N] = lam
    for i in range(N): D[i, i] = lam + 2 * mu
    for i in range(N, fl): D[i, i] = mu
    return D


def element_stiffness(N, deg, h, D):
    """K0 ((N (deg+1)^N)^2) of a full-density element with edge lengths h."""
    npe, fl = (deg + 1) ** N, (6 if N == 3 else 3)
    ke = N * npe
    gp, gw = gauss01(deg + 1)                      # degree 2 deg per axis -> deg + 1 points
    K = np.zeros((ke, ke))
    shear = np.array([1.0] * N + [2.0] * (fl - N))
    loc = list(itertools.product(range(deg + 1), repeat=N))     # local node -> per-axis index, row-major (last axis fastest)
    for q in itertools.product(range(deg + 1), repeat=N):
        w = np.prod([gw[k] for k in q])
        vd = [lagrange(deg, gp[q[d]]) for d in range(N)]
        B = np.zeros((ke, fl))
        for m, l in enumerate(loc):
            grad = np.array([np.prod([(vd[d][1][l[d]] / h[d]) if d == c else vd[d][0][l[d]] for d in range(N)]) for c in range(N)])
            for c in range(N):
                for i in range(N): B[N * m + c, sym_idx(N, c, i)] = 0.5 * grad[i]
                B[N * m + c, sym_idx(N, c, c)] = grad[c]
        K += w * (B * shear) @ D @ (B * shear).T
    return K * float(np.prod(h))


class Q2Sim:
    """TensorProductSimulator<double, 2, 2[, 2]> as far as the generic element path goes."""

    def __init__(self, ne, dmin, dmax, deg=2):
        self.ne = np.asarray(ne, dtype=np.int64); self.N = len(self.ne); self.deg = deg
        self.dmin, self.dmax = np.asarray(dmin, dtype=float), np.asarray(dmax, dtype=float)
        self.h = (self.dmax - self.dmin) / self.ne
        self.nn = deg * self.ne + 1
        self.num_nodes, self.num_elements = int(np.prod(self.nn)), int(np.prod(self.ne))
        self.law, self.E0, self.Emin, self.gamma, self.q = 0, 1.0, 1e-4, 3.0, 3.0
        self.rho = np.ones(self.num_elements)
        self.set_isotropic(1.0, 0.0)
        e = np.stack(np.meshgrid(*[np.arange(n) for n in self.ne], indexing="ij"), axis=-1).reshape(-1, self.N)
        loc = np.array(list(itertools.product(range(deg + 1), repeat=self.N)))
        self.enodes = np.ravel_multi_index(tuple((deg * e[:, None, :] + loc[None, :, :]).transpose(2, 0, 1)), tuple(self.nn))   # (ne, npe)

    def set_isotropic(self, young, nu): self.D = isotropic_tensor(self.N, young, nu); self.K0 = element_stiffness(self.N, self.deg, self.h, self.D)
    def set_interp(self, law, E0, Emin, gamma, q): self.law, self.E0, self.Emin, self.gamma, self.q = law, E0, Emin, gamma, q
    def set_densities(self, rho): self.rho = np.asarray(rho, dtype=float).copy()

    def E(self):                                    # m_updateYoungModuli (TensorProductSimulator.hh:2088-2102)
        r = self.rho
        return self.Emin + (r ** self.gamma if self.law == 0 else r / (1.0 + self.q * (1.0 - r))) * (self.E0 - self.Emin)

    def node_positions(self):
        g = np.meshgrid(*[self.dmin[d] + self.h[d] / self.deg * np.arange(self.nn[d]) for d in range(self.N)], indexing="ij")
        return np.stack([a.ravel() for a in g], axis=1)

    def apply_K(self, u):
        """f = K u  (element scatter, TPSStencils.hh:163-185); u, f: (numNodes, N)."""
        ue = u[self.enodes].reshape(self.num_elements, -1)                       # (ne, N npe), entry N m + c
        fe = (ue @ self.K0.T) * self.E()[:, None]
        f = np.zeros_like(u)
        np.add.at(f, self.enodes.ravel(), fe.reshape(-1, self.N))
        return f

    def element_energies(self, u):
        ue = u[self.enodes].reshape(self.num_elements, -1)
        return np.einsum("ei,ij,ej->e", ue, self.K0, ue)

    def assemble(self):
        import scipy.sparse as sp
        dofs = (self.N * self.enodes[:, :, None] + np.arange(self.N)).reshape(self.num_elements, -1)
        k = dofs.shape[1]
        rows, cols = np.repeat(dofs, k, axis=1).ravel(), np.tile(dofs, (1, k)).ravel()
        vals = (self.E()[:, None] * self.K0.ravel()[None, :]).ravel()
        n = self.N * self.num_nodes
        return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))

    def solve(self, f, fixed):
        """Direct solve of K u = f with the boolean (numNodes, N) mask `fixed` clamped to zero."""
        import scipy.sparse.linalg as spl
        K = self.assemble()
        free = np.flatnonzero(~np.asarray(fixed).ravel())
        u = np.zeros(self.N * self.num_nodes)
        u[free] = spl.spsolve(K[free][:, free].tocsc(), np.asarray(f).ravel()[free])
        return u.reshape(-1, self.N)
