"""Unit-cell stencil of a pair interaction: which basis atom, in which neighbouring cell, each symmetry
image of each input shell lands on.  This is the input of the on-device table builder
(asd_build_lattice_table); it restates the first half of the reference's `setup_nm`
(source/Hamiltonian/neighbourmap.f90:121-243: get_symops :361-516, get_fullnnlist :526-640, fold with
floor(x+5e-5) :199-201, match with squared distance < 0.01 :218) and the per-entry coupling conversion of
`setup_neighbour_hamiltonian` (source/Hamiltonian/hamiltonianinit.f90:1042,1064-1068).

Product code (host side, numpy); shares nothing with the test-only CPU checker.
"""
import numpy as np

TOL = 0.01  # neighbourmap.f90:90 (squared distance)


def symmetry_ops(sym):
    """Point-group operations in the reference's generation order (neighbourmap.f90:376-470)."""
    ops = []
    if sym == 0:
        ops.append(np.eye(3))
    elif sym == 1:
        for i in range(1, 4):
            for j in (0, 1):
                js = 1 if j == 0 else -1
                for x in (0, 1):
                    for y in (0, 1):
                        for z in (0, 1):
                            m = np.zeros((3, 3))
                            m[0, (i - js) % 3] = (-1.0) ** x      # Fortran mod(i-j_s,3)+1 -> 0-based column
                            m[1, i % 3] = (-1.0) ** y
                            m[2, (i + js) % 3] = (-1.0) ** z
                            ops.append(m)
    elif sym == 2:
        for j in (0, 1):
            for x in (0, 1):
                for y in (0, 1):
                    m = np.zeros((3, 3))
                    m[0, j % 2] = (-1.0) ** x
                    m[1, (j + 1) % 2] = (-1.0) ** y
                    m[2, 2] = 1.0
                    ops.append(m)
    elif sym == 3:
        half, rh = 0.5, np.sqrt(3.0) * 0.5
        for x in (0, 1):
            for y in (0, 1):
                for z in (0, 1):
                    ops.append(np.diag([(-1.0) ** x, (-1.0) ** y, (-1.0) ** z]))
        for x1 in (0, 1):
            for x2 in (0, 1):
                for y1 in (0, 1):
                    for y2 in (0, 1):
                        if (-1.0) ** (x1 + x2 + y1 + y2) < 0:
                            for z in (0, 1):
                                m = np.zeros((3, 3))
                                m[0, 0] = (-1.0) ** x1 * half
                                m[1, 0] = (-1.0) ** x2 * rh
                                m[0, 1] = (-1.0) ** y1 * rh
                                m[1, 1] = (-1.0) ** y2 * half
                                m[2, 2] = (-1.0) ** z
                                ops.append(m)
    else:
        raise ValueError('sym %d needs a sym.mat file' % sym)
    return ops


def _apply(op, v):
    # tvect(i) = sum_j redcoord(j) * sym_mats(i,j), accumulated j = 1..3 from zero like the reference
    out = np.zeros(3)
    for i in range(3):
        acc = 0.0
        for j in range(3):
            acc = acc + v[j] * op[i, j]
        out[i] = acc
    return out


def shell_images(redcoord, sym):
    """Distinct symmetry images of one shell vector, in generation order (get_fullnnlist)."""
    ops = symmetry_ops(sym)
    if len(ops) == 1:
        return [np.array(redcoord, dtype=float)]
    imgs = []
    for op in ops:
        t = _apply(op, redcoord)
        if all(((t - q) ** 2).sum() >= TOL for q in imgs):
            imgs.append(t)
    return imgs


def fold_basis(cell, bas):
    """geometry.f90:401-416: translate basis atoms into the first cell (floor(x + 1e-5))."""
    inv = np.linalg.inv(cell)  # rows of `cell` are C1,C2,C3: r = f @ cell  =>  f = r @ inv
    bas = np.array(bas, dtype=float).copy()
    for i0 in range(bas.shape[1]):
        f = bas[:, i0] @ inv
        bsf = np.floor(f + 1e-5)
        bas[:, i0] = bas[:, i0] - bsf[0] * cell[0] - bsf[1] * cell[1] - bsf[2] * cell[2]
    return bas


def stencil(cell, bas, atype, nn, redcoord, sym, nntype=None, ncell=None):
    """Returns nslot[NA], cell_atom[NA][maxslot] (1-based), cell_shift[NA][maxslot][3], shell_of[NA][maxslot].

    cell: 3x3 (rows C1,C2,C3); bas: (3,NA) folded basis; atype[NA] 1-based types; nn[NT] shells per type;
    redcoord[NT][maxshell][3]; nntype[NT][maxshell] (type the shell connects to) or None.
    Entry order = shell-major, then image order, then basis atom order (neighbourmap.f90:170-241).
    ncell = (N1, N2, N3): a SINGLE cell (N1*N2*N3 <= 1, e.g. a finite cluster given as one big cell) keeps no cell shifts at
    all -- nm_trunk is only filled for more than one cell and the hop is forced to zero (neighbourmap.f90:226-230,
    270-296) -- so a neighbour found through a folded basis position is kept whatever the boundary condition says.
    """
    cell = np.asarray(cell, dtype=float)
    det = np.linalg.det(cell)
    # explicit cofactor inverse in the reference's arrangement (invmatrix(r,c)), to fold exactly as it does
    C1, C2, C3 = cell
    inv = np.zeros((3, 3))
    inv[0, 0] = (C2[1] * C3[2] - C3[1] * C2[2]) / det
    inv[0, 1] = (C1[2] * C3[1] - C3[2] * C1[1]) / det
    inv[0, 2] = (C1[1] * C2[2] - C2[1] * C1[2]) / det
    inv[1, 0] = (C2[2] * C3[0] - C3[2] * C2[0]) / det
    inv[1, 1] = (C1[0] * C3[2] - C3[0] * C1[2]) / det
    inv[1, 2] = (C1[2] * C2[0] - C2[2] * C1[0]) / det
    inv[2, 0] = (C2[0] * C3[1] - C3[0] * C2[1]) / det
    inv[2, 1] = (C1[1] * C3[0] - C3[1] * C1[0]) / det
    inv[2, 2] = (C1[0] * C2[1] - C2[0] * C1[1]) / det
    na = bas.shape[1]
    entries = [[] for _ in range(na)]
    images = {}
    for i0 in range(na):
        it = int(atype[i0])
        for ish in range(int(nn[it - 1])):
            key = (it, ish)
            if key not in images:
                images[key] = shell_images(redcoord[it - 1][ish], sym)
            for v in images[key]:
                c = v + bas[:, i0]
                ic = np.array([c[0] * inv[0, k] + c[1] * inv[1, k] + c[2] * inv[2, k] for k in range(3)])
                bsf = np.floor(ic + 5.0e-5)
                r = c - bsf[0] * C1 - bsf[1] * C2 - bsf[2] * C3
                for ia in range(na):
                    if nntype is not None and int(atype[ia]) != int(nntype[it - 1][ish]):
                        continue
                    if ((r - bas[:, ia]) ** 2).sum() < TOL:
                        entries[i0].append((ia + 1, int(round(bsf[0])), int(round(bsf[1])), int(round(bsf[2])), ish))
    if ncell is not None and int(ncell[0]) * int(ncell[1]) * int(ncell[2]) <= 1:
        entries = [[(ja, 0, 0, 0, ish) for (ja, dx, dy, dz, ish) in x] for x in entries]
    maxslot = max(1, max(len(x) for x in entries))
    nslot = np.array([len(x) for x in entries], dtype=np.int32)
    cell_atom = np.ones((na, maxslot), dtype=np.int32)
    cell_shift = np.zeros((na, maxslot, 3), dtype=np.int32)
    shell_of = np.zeros((na, maxslot), dtype=np.int32)
    for i0 in range(na):
        for q, (ja, dx, dy, dz, ish) in enumerate(entries[i0]):
            cell_atom[i0, q] = ja
            cell_shift[i0, q] = (dx, dy, dz)
            shell_of[i0, q] = ish
    return nslot, cell_atom, cell_shift, shell_of


def couplings(nslot, cell_atom, shell_of, atype, xc, ammom, mry, mub, lexp=1):
    """ncoup of every stencil entry: xc * (2 mRy/mu_B) / m_i**lexp / m_j**lexp, zero when |m_i m_j| < 1e-6
    (hamiltonianinit.f90:1042,1062-1068).  xc[ncomp][NT][maxshell]."""
    xc = np.asarray(xc, dtype=float)
    if xc.ndim == 2:
        xc = xc[None]
    ncomp = xc.shape[0]
    na, maxslot = cell_atom.shape
    fc2 = 2.0 * mry / mub
    out = np.zeros((na, maxslot, ncomp))
    for i0 in range(na):
        for q in range(int(nslot[i0])):
            mi, mj = ammom[i0], ammom[cell_atom[i0, q] - 1]
            if abs(mi * mj) < float(np.float32(1e-6)):
                continue
            pi_, pj_ = (mi * mi, mj * mj) if lexp == 2 else (mi, mj)
            for a in range(ncomp):
                out[i0, q, a] = xc[a, atype[i0] - 1, shell_of[i0, q]] * fc2 / pi_ / pj_
    return out


def triangulation(n1, n2, n3, na):
    """The hard-coded triangulation of an a-priori triangular net (delaunay_tri_tri, source/Measurement/topology.f90:307-380):
    per cell (x, y, z) and basis atom two triangles, first all [0, +x, +y] then all [0, +y, +y - x], periodic in x and y.
    Returns simp(3, 2 * n1 * n2 * n3 * na), 1-based atom numbers in the reference's atom order."""
    z, y, x, it = np.meshgrid(np.arange(n3), np.arange(n2), np.arange(n1), np.arange(na), indexing='ij')
    z, y, x, it = z.ravel(), y.ravel(), x.ravel(), it.ravel()

    def atom(xx, yy):
        return na * (n1 * n2 * z + n1 * yy + xx) + it + 1

    xp, yp, xm = (x + 1) % n1, (y + 1) % n2, (x - 1) % n1
    first = np.stack([atom(x, y), atom(xp, y), atom(x, yp)])
    second = np.stack([atom(x, y), atom(x, yp), atom(xm, yp)])
    return np.asfortranarray(np.concatenate([first, second], axis=1), dtype=np.int32)
