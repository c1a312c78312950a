"""Gmsh 2.2 `.msh` files with fields: the on-disk format of the reference's density inputs (examples/densities/*.msh) and of its
field output (MeshFEM MSHFieldWriter / MSHFieldParser, included by TensorProductSimulator.hh:19-20; python bindings
3rdParty/MeshFEM/src/python_bindings/MSHFieldWriter_bindings.cc, MSHFieldParser_bindings.cc).  SURVEY.md section 8(f) rank 4.

* read_msh / write_msh: ASCII and binary (little-endian) files; $Nodes, $Elements, $NodeData, $ElementData.
* MSHFieldParser / MSHFieldWriter: the class surface the reference's python module exposes (vertices, elements, scalarField,
  vectorField, *FieldNames; addField).
* densities_from_msh / write_fields: the simulator side -- an element of the file belongs to the grid cell that contains its
  centroid (elementIndexFromMeshIO, TensorProductSimulator.hh:729-744), so the file's own element numbering is irrelevant;
  output uses getMesh's vertex and element ordering (:747-777).
Host-side only; nothing here is on the solve path."""
import struct

import numpy as np

# Gmsh element type -> nodes per element (line, triangle, quad, tet, hex, prism, pyramid, 2nd-order line / triangle / quad(9) / tet / hex(27))
_NODES = {1: 2, 2: 3, 3: 4, 4: 4, 5: 8, 6: 6, 7: 5, 8: 3, 9: 6, 10: 9, 11: 10, 12: 27, 15: 1, 16: 8, 17: 20}
_TYPE_OF = {(2, 3): 2, (2, 4): 3, (3, 4): 4, (3, 8): 5}     # (dimension, nodes per element) -> Gmsh type, first-order meshes


class _Reader:
    def __init__(self, data): self.b, self.i = data, 0
    def line(self):
        j = self.b.find(b"\n", self.i)
        if j < 0: j = len(self.b)
        s = self.b[self.i:j].decode("latin1").strip(); self.i = j + 1
        return s
    def skip_blank(self):
        while self.i < len(self.b) and self.b[self.i:self.i + 1] in (b"\n", b"\r", b" "): self.i += 1
    def take(self, n):
        s = self.b[self.i:self.i + n]
        if len(s) != n: raise RuntimeError("msh: truncated binary section")
        self.i += n
        return s
    def eof(self): return self.i >= len(self.b)


def read_msh(path):
    """-> dict(vertices (nv, 3), elements (ne, k) zero-based, element_type, binary, fields {name: (domain, array (n, ncomp))})
    with domain 'node' / 'element'.  Node and element ids are renumbered to positions 0..n-1 in file order (MeshIO_MSH)."""
    r = _Reader(open(path, "rb").read())
    out = dict(vertices=None, elements=None, element_type=None, binary=False, fields={})
    node_pos, elem_pos = {}, {}
    while not r.eof():
        r.skip_blank()
        if r.eof(): break
        sec = r.line()
        if sec == "$MeshFormat":
            ver, ftype, dsize = r.line().split()
            if not ver.startswith("2"): raise RuntimeError("msh: only format version 2.x is supported (found %s)" % ver)
            if int(dsize) != 8: raise RuntimeError("msh: data size must be 8")
            out["binary"] = int(ftype) == 1
            if out["binary"]:
                if struct.unpack("<i", r.take(4))[0] != 1: raise RuntimeError("msh: big-endian binary files are not supported")
            r.skip_blank(); r.line()
        elif sec == "$Nodes":
            n = int(r.line())
            if out["binary"]:
                rec = np.frombuffer(r.take(28 * n), dtype=np.dtype([("id", "<i4"), ("p", "<f8", 3)]))
                ids, V = rec["id"].astype(np.int64), rec["p"].copy()
            else:
                rows = np.array([r.line().split() for _ in range(n)], dtype=np.float64).reshape(n, 4)
                ids, V = rows[:, 0].astype(np.int64), rows[:, 1:4].copy()
            node_pos = {int(k): i for i, k in enumerate(ids)}
            out["vertices"] = V
            r.skip_blank(); r.line()
        elif sec == "$Elements":
            n = int(r.line())
            els, ids, etype = [], [], None
            if out["binary"]:
                got = 0
                while got < n:
                    t, cnt, ntags = struct.unpack("<3i", r.take(12))
                    k = _NODES[t]
                    rec = np.frombuffer(r.take(4 * (1 + ntags + k) * cnt), dtype="<i4").reshape(cnt, 1 + ntags + k)
                    if etype is None or _NODES[t] > _NODES[etype]: etype = t
                    for row in rec: els.append((t, row[1 + ntags:])); ids.append(int(row[0]))
                    got += cnt
            else:
                for _ in range(n):
                    f = [int(v) for v in r.line().split()]
                    t, ntags = f[1], f[2]
                    els.append((t, np.array(f[3 + ntags:]))); ids.append(f[0])
                    if etype is None or _NODES[t] > _NODES[etype]: etype = t
            # keep the elements of the highest-dimensional type only (MeshIO drops lower-dimensional boundary elements)
            keep = [(i, e) for i, (t, e) in zip(ids, els) if t == etype]
            elem_pos = {i: k for k, (i, _) in enumerate(keep)}
            out["elements"] = np.array([[node_pos[int(v)] for v in e] for _, e in keep], dtype=np.int64)
            out["element_type"] = etype
            r.skip_blank(); r.line()
        elif sec in ("$NodeData", "$ElementData"):
            stags = [r.line().strip('"') for _ in range(int(r.line()))]
            for _ in range(int(r.line())): r.line()
            itags = [int(r.line()) for _ in range(int(r.line()))]
            ncomp, nvals = itags[1], itags[2]
            if out["binary"]:
                rec = np.frombuffer(r.take((4 + 8 * ncomp) * nvals), dtype=np.dtype([("id", "<i4"), ("v", "<f8", ncomp)]))
                ids, vals = rec["id"].astype(np.int64), rec["v"].reshape(nvals, ncomp).copy()
            else:
                rows = np.array([r.line().split() for _ in range(nvals)], dtype=np.float64).reshape(nvals, 1 + ncomp)
                ids, vals = rows[:, 0].astype(np.int64), rows[:, 1:].copy()
            pos = node_pos if sec == "$NodeData" else elem_pos
            size = len(pos)
            arr = np.zeros((size, ncomp))
            idx = np.array([pos.get(int(k), -1) for k in ids])
            arr[idx[idx >= 0]] = vals[idx >= 0]
            out["fields"][stags[0] if stags else "field%d" % len(out["fields"])] = ("node" if sec == "$NodeData" else "element", arr)
            r.skip_blank(); r.line()
        else:                                   # unknown section: skip to its end marker
            end = "$End" + sec[1:]
            while not r.eof() and r.line() != end: pass
    return out


def write_msh(path, V, F, fields=None, binary=True):
    """V (nv, 2 or 3), F (ne, k) zero-based with Gmsh node ordering, fields {name: (domain, array)} (domain 'node' / 'element')."""
    V = np.asarray(V, dtype=np.float64); F = np.asarray(F, dtype=np.int64)
    P = np.zeros((V.shape[0], 3)); P[:, :V.shape[1]] = V
    k = F.shape[1]
    dim = 2 if k == 3 or (k == 4 and not np.any(P[:, 2])) else 3       # 4 nodes: a quad of a planar mesh, else a tetrahedron
    if (dim, k) not in _TYPE_OF: raise RuntimeError("msh: unsupported element with %d nodes in %dD" % (k, dim))
    etype = _TYPE_OF[(dim, F.shape[1])]
    with open(path, "wb") as f:
        w = lambda s: f.write(s.encode("latin1"))
        w("$MeshFormat\n2.2 %d 8\n" % (1 if binary else 0))
        if binary: f.write(struct.pack("<i", 1)); w("\n")
        w("$EndMeshFormat\n$Nodes\n%d\n" % P.shape[0])
        if binary:
            rec = np.zeros(P.shape[0], dtype=np.dtype([("id", "<i4"), ("p", "<f8", 3)]))
            rec["id"] = np.arange(1, P.shape[0] + 1); rec["p"] = P
            f.write(rec.tobytes()); w("\n")
        else:
            for i, p in enumerate(P): w("%d %.17g %.17g %.17g\n" % (i + 1, p[0], p[1], p[2]))
        w("$EndNodes\n$Elements\n%d\n" % F.shape[0])
        if binary:
            f.write(struct.pack("<3i", etype, F.shape[0], 0))
            rec = np.empty((F.shape[0], 1 + F.shape[1]), dtype="<i4")
            rec[:, 0] = np.arange(1, F.shape[0] + 1); rec[:, 1:] = F + 1
            f.write(rec.tobytes()); w("\n")
        else:
            for i, e in enumerate(F): w("%d %d 0 %s\n" % (i + 1, etype, " ".join(str(int(v) + 1) for v in e)))
        w("$EndElements\n")
        for name, (domain, arr) in (fields or {}).items():
            a = np.asarray(arr, dtype=np.float64)
            if a.ndim == 1: a = a[:, None]
            n = P.shape[0] if domain == "node" else F.shape[0]
            if a.shape[0] != n: raise RuntimeError("msh: field '%s' has %d rows, expected %d" % (name, a.shape[0], n))
            ncomp = a.shape[1]
            if ncomp == 2:                      # vectors are written with 3 components (Gmsh knows 1, 3 and 9)
                a = np.hstack([a, np.zeros((n, 1))]); ncomp = 3
            sec = "NodeData" if domain == "node" else "ElementData"
            w("$%s\n1\n\"%s\"\n0\n3\n0\n%d\n%d\n" % (sec, name, ncomp, n))
            if binary:
                rec = np.zeros(n, dtype=np.dtype([("id", "<i4"), ("v", "<f8", ncomp)]))
                rec["id"] = np.arange(1, n + 1); rec["v"] = a
                f.write(rec.tobytes()); w("\n")
            else:
                for i, row in enumerate(a): w("%d %s\n" % (i + 1, " ".join("%.17g" % v for v in row)))
            w("$End%s\n" % sec)


class DomainType:
    PER_ELEMENT, PER_NODE, GUESS, ANY, UNKNOWN = "element", "node", "guess", "any", "unknown"


class MSHFieldParser:
    """MSHFieldParser (MSHFieldParser_bindings.cc:8-30): vertices(), elements(), scalarField(name), vectorField(name), field names."""

    def __init__(self, mshPath, permitDimMismatch=True):
        self._m = read_msh(mshPath)
    def vertices(self): return self._m["vertices"].copy()
    def elements(self): return self._m["elements"].copy()
    def numVertices(self): return int(self._m["vertices"].shape[0])
    def numElements(self): return int(self._m["elements"].shape[0])
    def meshDimension(self): return 2 if self._m["element_type"] in (2, 3, 9, 10, 16) else 3
    def meshDegree(self): return 1 if self._m["element_type"] in (2, 3, 4, 5) else 2
    def _names(self, pred, domainType):
        return [k for k, (d, a) in self._m["fields"].items() if pred(a) and domainType in (DomainType.ANY, DomainType.GUESS, d)]
    def scalarFieldNames(self, domainType=DomainType.ANY): return self._names(lambda a: a.shape[1] == 1, domainType)
    def vectorFieldNames(self, domainType=DomainType.ANY): return self._names(lambda a: a.shape[1] == 3, domainType)
    def symmetricMatrixFieldNames(self, domainType=DomainType.ANY): return self._names(lambda a: a.shape[1] == 9, domainType)
    def _field(self, name, domainType):
        if name not in self._m["fields"]: raise RuntimeError("Field '%s' not found" % name)
        d, a = self._m["fields"][name]
        if domainType not in (DomainType.ANY, DomainType.GUESS, d): raise RuntimeError("Field '%s' has domain type %s" % (name, d))
        return a
    def scalarField(self, name, domainType=DomainType.ANY): return self._field(name, domainType)[:, 0].copy()
    def vectorField(self, name, domainType=DomainType.ANY):
        a = self._field(name, domainType)
        return a[:, :self.meshDimension()].copy()


class MSHFieldWriter:
    """MSHFieldWriter(path, V, F, binary=True).addField(name, field, domainType) (MSHFieldWriter_bindings.cc:17-45); the file is
    written when the writer is closed or garbage-collected."""

    def __init__(self, path, V, F, binary=True):
        self._path, self._V, self._F, self._binary, self._fields, self._open = path, np.asarray(V), np.asarray(F), binary, {}, True
    def addField(self, name, field, domainType=DomainType.GUESS):
        a = np.asarray(field, dtype=np.float64)
        if a.ndim == 1: a = a[:, None]
        if domainType in (DomainType.GUESS, DomainType.ANY):
            nv, ne = self._V.shape[0], self._F.shape[0]
            if a.shape[0] == ne and a.shape[0] != nv: domainType = DomainType.PER_ELEMENT
            elif a.shape[0] == nv: domainType = DomainType.PER_NODE
            else: raise RuntimeError("Cannot guess the domain type of field '%s'" % name)
        self._fields[name] = (domainType, a)
    def close(self):
        if self._open:
            write_msh(self._path, self._V, self._F, self._fields, self._binary); self._open = False
    def __del__(self):
        try: self.close()
        except Exception: pass
    def __enter__(self): return self
    def __exit__(self, *a): self.close()


def element_indices_from_msh(tps, vertices, elements, bbox=None):
    """Grid cell (flat element index of the simulator) of every element of a mesh: the cell that contains the element's centroid,
    with the mesh's bounding box mapped onto the simulator's domain (elementIndexFromMeshIO, TensorProductSimulator.hh:729-744)."""
    ne = np.asarray(tps.NbElementsPerDimension, dtype=np.int64)
    N = len(ne)
    V = np.asarray(vertices, dtype=np.float64)
    c = V[np.asarray(elements)].mean(axis=1)[:, :N]
    lo, hi = (V.min(axis=0)[:N], V.max(axis=0)[:N]) if bbox is None else (np.asarray(bbox[0])[:N], np.asarray(bbox[1])[:N])
    t = (c - lo) / (hi - lo)                                   # interpolation coordinates in the mesh's bounding box
    cell = np.minimum(np.floor(t * ne).astype(np.int64), ne - 1)
    if np.any(cell < 0): raise RuntimeError("msh: element outside the grid")
    return np.ravel_multi_index(tuple(cell.T), tuple(ne))


def densities_from_msh(tps, path, field="density"):
    """The per-element scalar field `field` of a .msh file as the simulator's flat density array (examples/densities/*.msh)."""
    m = read_msh(path)
    if field not in m["fields"] or m["fields"][field][0] != "element": raise RuntimeError("msh: no per-element field '%s'" % field)
    idx = element_indices_from_msh(tps, m["vertices"], m["elements"])
    if len(idx) != int(np.prod(np.asarray(tps.NbElementsPerDimension))) or len(np.unique(idx)) != len(idx):
        raise RuntimeError("msh: the mesh's elements do not tile the simulator's grid one to one")
    rho = np.zeros(len(idx)); rho[idx] = m["fields"][field][1][:, 0]
    return rho


def write_fields(tps, path, element_fields=None, node_fields=None, binary=True):
    """The simulator's mesh (getMesh ordering) with per-element / per-node fields, e.g. densities and displacements."""
    V, F = tps.getMesh()
    N = len(np.asarray(tps.NbElementsPerDimension))
    fields = {k: ("element", v) for k, v in (element_fields or {}).items()}
    fields.update({k: ("node", v) for k, v in (node_fields or {}).items()})
    write_msh(path, V[:, :3] if N == 3 else V[:, :2], F, fields, binary)
