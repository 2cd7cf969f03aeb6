"""TEST INFRASTRUCTURE ONLY -- minimal readers for the reference's on-disk input formats.

Restates just enough of the reference parsers to load its regression fixtures:
  inpsd.dat keywords   source/Input/inputhandler.f90:52 (read_parameters; case-insensitive `keyword value`)
  posfile              source/Input/inputhandler_ext.f90:59-128   (read_positions)
  momfile              source/Input/inputhandler_ext.f90:228-330  (read_moments, set_landeg=0)
  jfile                source/Input/inputhandler_ext.f90:444-627  (read_exchange) + :1007-1042 (getNeighVec)
  dmfile               source/Input/inputhandler_ext.f90:1095-1190 (read_dmdata)
  bqfile               source/Input/inputhandler_ext.f90:2023-2146 (read_bqdata)
  kfile                source/Input/inputhandler_ext.f90:1069-1089 (read_anisotropy)
Defaults follow source/Input/inputdata.f90:300-530.

Used by tests/golden/make_fixtures.py (run in the build container, where /root/reference exists) to turn a
reference fixture directory into a self-contained JSON fixture; nothing here runs on the GPU box against
/root/reference.
"""
import os
import numpy as np


def _f(tok):
    return float(tok.replace('d', 'e').replace('D', 'e'))


def _rows(path):
    out = []
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t or t[0].startswith('#'):
                continue
            out.append(t)
    return out


DEFAULTS = dict(
    simid='_UppASD_', ncell=(1, 1, 1), bc=('0', '0', '0'), cell=np.eye(3), sym=0, posfiletype='C', maptype=1,
    mensemble=1, tseed=1, sdealgh=1, initmag=3, mode='S', temp=0.0, nstep=1, damping=0.05, timestep=1.0e-16,
    hfield=(0.0, 0.0, 0.0), do_reduced='N', do_sortcoup='N', mompar=0, landeg_glob=2.0, do_dm=0, do_bq=0, do_jtensor=0,
    do_anisotropy=0, mcnstep=0, avrg_step=100, cumu_step=50, cumu_buff=10, do_avrg='N', do_cumu='N',
    plotenergy=0, map_multiple=False, gpu_mode=0, ip_mode='N', aunits='N', do_ralloy=0,
)


def read_inpsd(path):
    """Parse the keywords this build understands; unknown keywords are ignored (like a wide `select case`)."""
    d = dict(DEFAULTS)
    base = os.path.dirname(os.path.abspath(path))
    lines = open(path).read().splitlines()
    i = 0
    files = {}
    while i < len(lines):
        t = lines[i].split()
        i += 1
        if not t:
            continue
        key = t[0].lower()
        v = t[1:]
        if key == 'simid':
            d['simid'] = v[0][:8]                      # character(len=8) :: simid
        elif key == 'ncell':
            d['ncell'] = tuple(int(x) for x in v[:3])
        elif key == 'bc':
            d['bc'] = tuple(x.upper() for x in v[:3])
        elif key == 'cell':
            rows = [v[:3]]
            while len(rows) < 3:
                rows.append(lines[i].split()[:3])
                i += 1
            d['cell'] = np.array([[_f(x) for x in r] for r in rows])
        elif key == 'sym':
            d['sym'] = int(v[0])
        elif key in ('posfile', 'momfile', 'exchange', 'dm', 'bq', 'anisotropy'):
            files[key] = os.path.normpath(os.path.join(base, v[0]))
            if key == 'dm':
                d['do_dm'] = 1
            if key == 'bq':
                d['do_bq'] = 1
            if key == 'anisotropy':
                d['do_anisotropy'] = 1
        elif key == 'posfiletype':
            d['posfiletype'] = v[0].upper()
        elif key == 'maptype':
            d['maptype'] = int(v[0])
        elif key == 'mensemble':
            d['mensemble'] = int(v[0])
        elif key == 'tseed':
            d['tseed'] = int(v[0])
        elif key == 'do_ralloy':
            d['do_ralloy'] = int(v[0])
        elif key == 'sdealgh':
            d['sdealgh'] = int(v[0])
        elif key == 'initmag':
            d['initmag'] = int(v[0])
        elif key == 'mode':
            d['mode'] = v[0].upper()
        elif key == 'temp':
            d['temp'] = _f(v[0])
        elif key == 'nstep':
            d['nstep'] = int(v[0])
        elif key == 'mcnstep':
            d['mcnstep'] = int(v[0])
        elif key == 'damping':
            d['damping'] = _f(v[0])
        elif key == 'timestep':
            d['timestep'] = _f(v[0])
        elif key == 'hfield':
            d['hfield'] = tuple(_f(x) for x in v[:3])
        elif key == 'do_jtensor':
            d['do_jtensor'] = int(v[0])
        elif key == 'do_reduced':
            d['do_reduced'] = v[0].upper()
        elif key == 'do_sortcoup':
            d['do_sortcoup'] = v[0].upper()
        elif key == 'mompar':
            d['mompar'] = int(v[0])
        elif key == 'avrg_step':
            d['avrg_step'] = int(v[0])
        elif key == 'cumu_step':
            d['cumu_step'] = int(v[0])
        elif key == 'cumu_buff':
            d['cumu_buff'] = int(v[0])
        elif key == 'do_avrg':
            d['do_avrg'] = v[0].upper()
        elif key == 'do_cumu':
            d['do_cumu'] = v[0].upper()
        elif key == 'plotenergy':
            d['plotenergy'] = int(v[0])
        elif key == 'gpu_mode':
            d['gpu_mode'] = int(v[0])
        elif key == 'ip_mode':
            d['ip_mode'] = v[0].upper()
        elif key == 'aunits':
            d['aunits'] = v[0].upper()
    d['files'] = files
    return d


def read_positions(path, cell, posfiletype='C'):
    rows = _rows(path) if isinstance(path, str) else path
    na = max(int(r[0]) for r in rows)
    bas = np.zeros((3, na), order='F')
    atype = np.zeros(na, dtype=np.int32)
    for r in rows[:na]:
        isite, itype = int(r[0]), int(r[1])
        p = np.array([_f(x) for x in r[2:5]])
        if posfiletype == 'D':
            p = p[0] * cell[0] + p[1] * cell[1] + p[2] * cell[2]
        bas[:, isite - 1] = p
        atype[isite - 1] = itype
    return bas, atype


def read_moments(path, na, landeg_glob=2.0):
    ammom = np.zeros(na)
    aemom = np.zeros((3, na), order='F')
    for r in (_rows(path) if isinstance(path, str) else path):
        isite = int(r[0])
        ammom[isite - 1] = _f(r[2])
        e = np.array([_f(x) for x in r[3:6]])
        # aemom_tmp = norm2(aemom_inp); aemom_inp = aemom_inp / aemom_tmp
        nrm = np.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
        aemom[:, isite - 1] = e / nrm
    landeg = np.full(na, landeg_glob)
    return ammom, aemom, landeg


def _neigh_vec(r_tmp, isite, jsite, bas, cell, maptype, posfiletype):
    if maptype == 2:
        return np.array([bas[a, jsite - 1] - bas[a, isite - 1] + cell[0][a] * r_tmp[0] + cell[1][a] * r_tmp[1]
                         + cell[2][a] * r_tmp[2] for a in range(3)])
    if posfiletype == 'D':
        return np.array([r_tmp[0] * cell[0][a] + r_tmp[1] * cell[1][a] + r_tmp[2] * cell[2][a] for a in range(3)])
    return np.array(r_tmp, dtype=float)


def read_pair_file(path, nt, atype_inp, bas, cell, maptype, posfiletype, ncomp, with_nntype):
    """Shell table builder shared by read_exchange / read_dmdata / read_bqdata (do_ralloy=0).

    Returns nn(NT), redcoord(NT,maxshells,3), xc(ncomp,NT,maxshells), nntype(NT,maxshells) or None."""
    rows = _rows(path) if isinstance(path, str) else path
    tol = 1.0e-5
    nn = np.zeros(nt, dtype=np.int32)
    red = [[] for _ in range(nt)]
    val = [[] for _ in range(nt)]
    ntyp = [[] for _ in range(nt)]
    for r in rows:
        isite, jsite = int(r[0]), int(r[1])
        r_tmp = [_f(x) for x in r[2:5]]
        v = [_f(x) for x in r[5:5 + ncomp]]
        itype = int(atype_inp[isite - 1])
        r_red = _neigh_vec(r_tmp, isite, jsite, bas, cell, maptype, posfiletype)
        unique = True
        for ish in range(nn[itype - 1]):
            c = red[itype - 1][ish]
            norm = (r_red[0] - c[0]) ** 2 + (r_red[1] - c[1]) ** 2 + (r_red[2] - c[2]) ** 2
            if norm < tol:
                unique = False
                val[itype - 1][ish] = v
        if unique:
            nn[itype - 1] += 1
            red[itype - 1].append(r_red)
            val[itype - 1].append(v)
            ntyp[itype - 1].append(int(atype_inp[jsite - 1]))
    ms = int(nn.max())
    redcoord = np.zeros((nt, ms, 3), order='F')
    xc = np.zeros((ncomp, nt, ms), order='F')
    nntype = np.zeros((nt, ms), dtype=np.int32, order='F')
    for t in range(nt):
        for s in range(nn[t]):
            redcoord[t, s, :] = red[t][s]
            xc[:, t, s] = val[t][s]
            nntype[t, s] = ntyp[t][s]
    return nn, redcoord, xc, (nntype if with_nntype else None)


def read_tensor_file(path, nt, atype_inp, bas, cell, maptype, posfiletype):
    """jfile in tensor format (do_jtensor 1; inputhandler_ext.f90:658-740, read_exchange_tensor_base): nine numbers per
    line read into j_tmp(3,3) in Fortran order and then TRANSPOSED (:686-687), i.e. the file holds the tensor row by
    row; neighbour type is not distinguished (jtype = 1, :704).  Returns nn, redcoord, xc(9, NT, shells) with xc the
    column-major flattening of J(a,b), no nntype."""
    nn, red, xc, _ = read_pair_file(path, nt, atype_inp, bas, cell, maptype, posfiletype, 9, False)
    out = np.zeros_like(xc)
    for a in range(3):
        for b in range(3):
            out[a + 3 * b] = xc[3 * a + b]          # J(a,b) = file[3a + b], stored at Fortran offset a + 3b
    return nn, red, out, None


def read_anisotropy(path, na):
    atyp = np.zeros(na, dtype=np.int32)
    an = np.zeros((na, 6), order='F')
    for r in (_rows(path) if isinstance(path, str) else path)[:na]:
        iat = int(r[0])
        atyp[iat - 1] = int(r[1])
        an[iat - 1, :] = [_f(x) for x in r[2:8]]
    return atyp, an


def load_fixture(fx):
    """fx: dict with 'inp' (read_inpsd output, JSON-ified) and raw token rows 'posfile','momfile','jfile'
    [, 'dmfile','bqfile','kfile'].  Returns the argument tuple for oracle.orc.build_system."""
    inp = dict(fx['inp'])
    inp['cell'] = np.array(inp['cell'], dtype=float)
    inp['ncell'] = tuple(inp['ncell'])
    inp['bc'] = tuple(inp['bc'])
    inp['hfield'] = tuple(inp['hfield'])
    bas, atype_inp = read_positions(fx['posfile'], inp['cell'], inp['posfiletype'])
    na = bas.shape[1]
    nt = int(atype_inp.max())
    ammom, aemom, landeg = read_moments(fx['momfile'], na, inp['landeg_glob'])

    def mk(key, ncomp, with_nntype):
        if key not in fx or fx[key] is None:
            return None
        return lambda S: read_pair_file(fx[key], nt, atype_inp, S['bas'], inp['cell'], inp['maptype'],
                                        inp['posfiletype'], ncomp, with_nntype)
    aniso = read_anisotropy(fx['kfile'], na) if fx.get('kfile') else None
    ex = mk('jfile', 1, True)
    if inp.get('do_jtensor', 0) == 1:
        ex = lambda S: read_tensor_file(fx['jfile'], nt, atype_inp, S['bas'], inp['cell'], inp['maptype'], inp['posfiletype'])
    return inp, bas, atype_inp, ammom, aemom, landeg, ex, mk('dmfile', 3, False), mk('bqfile', 1, False), aniso


# ---- random alloys (do_ralloy 1) ------------------------------------------------------------------------------------
def read_positions_alloy(path, cell, posfiletype='C'):
    """read_positions_alloy (inputhandler_ext.f90:139-216): rows `site type chem conc x y z`.
    Returns bas(3,NA), atype_inp(NA), nch(NA), chconc(NA,Nchmax)."""
    rows = _rows(path) if isinstance(path, str) else path
    na = max(int(r[0]) for r in rows)
    nchmax = max(int(r[2]) for r in rows)
    bas = np.zeros((3, na), order='F')
    atype = np.zeros(na, dtype=np.int32)
    nch = np.zeros(na, dtype=np.int32)
    chconc = np.zeros((na, nchmax), order='F')
    for r in rows:
        isite, itype, ichem = int(r[0]), int(r[1]), int(r[2])
        p = np.array([_f(x) for x in r[4:7]])
        if posfiletype == 'D':
            p = p[0] * cell[0] + p[1] * cell[1] + p[2] * cell[2]
        bas[:, isite - 1] = p
        nch[isite - 1] = max(nch[isite - 1], ichem)
        atype[isite - 1] = itype
        chconc[isite - 1, ichem - 1] = _f(r[3])
    return bas, atype, nch, chconc


def read_moments_alloy(path, na, nchmax, landeg_glob=2.0):
    """read_moments (inputhandler_ext.f90:228-321), set_landeg 0, no LSF / induced moments: rows `site chem mom ex ey ez`.
    Returns ammom_inp(NA,Nchmax), aemom_inp(3,NA,Nchmax), Landeg_ch(NA,Nchmax)."""
    ammom = np.zeros((na, nchmax), order='F')
    aemom = np.zeros((3, na, nchmax), order='F')
    for r in (_rows(path) if isinstance(path, str) else path):
        isite, ichem = int(r[0]), int(r[1])
        ammom[isite - 1, ichem - 1] = _f(r[2])
        e = np.array([_f(x) for x in r[3:6]])
        nrm = np.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
        aemom[:, isite - 1, ichem - 1] = e / nrm
    return ammom, aemom, np.full((na, nchmax), landeg_glob, order='F')


def read_pair_file_alloy(path, nt, nchmax, atype_inp, bas, cell, maptype, posfiletype):
    """read_exchange with do_ralloy 1 (inputhandler_ext.f90:444-627, the `else` branches at :487-531): rows
    `isite jsite ichem jchem r1 r2 r3 J`; the shell is identified by the vector as for do_ralloy 0, the coupling is filed under
    jc(itype, shell, ichem, jchem).  Returns nn(NT), redcoord(NT,ms,3), xc(1,NT,ms,Nchmax,Nchmax), nntype(NT,ms)."""
    rows = _rows(path) if isinstance(path, str) else path
    tol = 1.0e-5
    nn = np.zeros(nt, dtype=np.int32)
    red = [[] for _ in range(nt)]
    val = [[] for _ in range(nt)]
    ntyp = [[] for _ in range(nt)]
    for r in rows:
        isite, jsite, ichem, jchem = int(r[0]), int(r[1]), int(r[2]), int(r[3])
        r_tmp = [_f(x) for x in r[4:7]]
        j_tmp = _f(r[7])
        itype = int(atype_inp[isite - 1])
        r_red = _neigh_vec(r_tmp, isite, jsite, bas, cell, maptype, posfiletype)
        unique = True
        for ish in range(nn[itype - 1]):
            c = red[itype - 1][ish]
            norm = (r_red[0] - c[0]) ** 2 + (r_red[1] - c[1]) ** 2 + (r_red[2] - c[2]) ** 2
            if norm < tol:
                unique = False
                val[itype - 1][ish][ichem - 1, jchem - 1] = j_tmp
        if unique:
            nn[itype - 1] += 1
            red[itype - 1].append(r_red)
            m = np.zeros((nchmax, nchmax))
            m[ichem - 1, jchem - 1] = j_tmp
            val[itype - 1].append(m)
            ntyp[itype - 1].append(int(atype_inp[jsite - 1]))
    ms = int(nn.max())
    redcoord = np.zeros((nt, ms, 3), order='F')
    xc = np.zeros((1, nt, ms, nchmax, nchmax), order='F')
    nntype = np.zeros((nt, ms), dtype=np.int32, order='F')
    for t in range(nt):
        for s in range(nn[t]):
            redcoord[t, s, :] = red[t][s]
            xc[0, t, s] = val[t][s]
            nntype[t, s] = ntyp[t][s]
    return nn, redcoord, xc, nntype


def load_alloy_fixture(fx):
    """argument tuple of oracle.orc.build_alloy_system from a fixture with raw token rows (posfile, momfile, jfile)"""
    inp = dict(fx['inp'])
    inp['cell'] = np.array(inp['cell'], dtype=float)
    inp['ncell'] = tuple(inp['ncell'])
    inp['bc'] = tuple(inp['bc'])
    inp['hfield'] = tuple(inp['hfield'])
    bas, atype_inp, nch, chconc = read_positions_alloy(fx['posfile'], inp['cell'], inp['posfiletype'])
    na, nchmax = chconc.shape
    nt = int(atype_inp.max())
    ammom, aemom, landeg = read_moments_alloy(fx['momfile'], na, nchmax, inp['landeg_glob'])
    ex = lambda S: read_pair_file_alloy(fx['jfile'], nt, nchmax, atype_inp, S['bas'], inp['cell'], inp['maptype'], inp['posfiletype'])
    return inp, bas, atype_inp, nch, chconc, ammom, aemom, landeg, ex
