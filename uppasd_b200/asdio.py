"""On-disk formats of the reference, for the slice of it this path serves.

Readers (what the driver needs to set a simulation up from a reference run directory):
  inpsd.dat            `keyword value` lines, case-insensitive, `d` exponents, block keywords (cell, ip_nphase,
                       ip_mcanneal, ntraj) -- source/Input/inputhandler.f90:52 ff.; defaults inputdata.f90:300-530,
                       prn_averages.f90:254-263, prn_trajectories.f90:112-125
  posfile / momfile    source/Input/inputhandler_ext.f90:59-128, 228-330
  jfile / dmfile / bqfile   :444-627, 1095-1190, 2023-2146 with getNeighVec :1007-1042 (maptype 1 / 2, posfiletype C / D)
  kfile                :1069-1089
  restart file         source/System/restart.f90:320-381 (new format written by prn_mag_conf_iter :186-246)
Writers (what the reference's own regression YAMLs read back, tests/bergtest.py + extractoutput.py):
  averages.<simid>.out, cumulants.<simid>.out, totenergy.<simid>.out, trajectory.<simid>.<atom>.<ens>.out,
  restart.<simid>.out, coord.<simid>.out
with the reference's Fortran edit descriptors (i8, es16.8, ...) reproduced digit for digit.

Host-side product code (numpy only).
"""
import os

import numpy as np


class InputError(ValueError):
    pass


def _num(tok):
    t = tok.strip().rstrip(',')
    return float(t.replace('d', 'e').replace('D', 'e'))


def _flag(tok):
    t = tok.strip().strip('.').upper()
    return t[0] if t else 'N'


def _data_rows(path):
    rows = []
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if t and not t[0].startswith(('#', '!')):
                rows.append(t)
    return rows


# ---------------------------------------------------------------------------------------------------------------
# inpsd.dat
# ---------------------------------------------------------------------------------------------------------------
def defaults():
    return dict(
        simid='_UppASD_', ncell=(1, 1, 1), bc=('0', '0', '0'), cell=np.eye(3), sym=0, alat=1.0, aunits='N',
        posfile=None, momfile=None, exchange=None, dm=None, bq=None, anisotropy=None, restartfile=None,
        posfiletype='C', maptype=1, do_ralloy=0, mensemble=1, tseed=1, sdealgh=1, ipsdealgh=-1, initmag=3, mode='S',
        ip_mode='N', temp=0.0, nstep=1, mcnstep=0, damping=0.05, timestep=1.0e-16, hfield=(0.0, 0.0, 0.0),
        ip_hfield=(0.0, 0.0, 0.0), ip_temp=0.0, ip_nphase=[], ip_mcanneal=[], ip_mcnstep=0, do_reduced='N', do_sortcoup='N',
        mompar=0, landeg_glob=2.0, do_avrg='Y', avrg_step=100, avrg_buff=10, do_cumu='N', cumu_step=50, cumu_buff=10,
        plotenergy=0, do_tottraj='N', tottraj_step=1000, tottraj_buff=10, trajectories=[], do_prnstruct=0,
        gpu_mode=0, gpu_rng_seed=0, do_jtensor=0, do_bpulse=0, bpulsefile=None, map_multiple=False, relaxed_if=False, do_proj_avrg='N', do_cumu_proj='N', skyno='N',
        skyno_step=100, skyno_buff=10)


_SCALAR_INT = {'sym', 'maptype', 'do_ralloy', 'mensemble', 'tseed', 'sdealgh', 'ipsdealgh', 'initmag', 'nstep', 'mcnstep',
               'mompar', 'avrg_step', 'avrg_buff', 'cumu_step', 'cumu_buff', 'plotenergy', 'tottraj_step', 'tottraj_buff',
               'do_prnstruct', 'gpu_mode', 'gpu_rng_seed', 'do_jtensor', 'ip_mcnstep', 'skyno_step', 'skyno_buff', 'do_bpulse'}
_SCALAR_REAL = {'alat', 'temp', 'damping', 'timestep', 'ip_temp'}
_SCALAR_FLAG = {'aunits', 'posfiletype', 'do_reduced', 'do_sortcoup', 'do_avrg', 'do_cumu', 'do_tottraj', 'mode', 'ip_mode', 'do_proj_avrg', 'do_cumu_proj', 'skyno'}
_FILES = {'posfile', 'momfile', 'exchange', 'dm', 'bq', 'anisotropy', 'restartfile', 'bpulsefile'}


# Reference keywords that change the Hamiltonian, the dynamics or what a measurement means, and that this path does NOT serve
# (inputhandler.f90's `select case`; the reading modules of prn_*, sLLG, fieldpulse, temperature).  keyword -> the values that
# leave the feature off.  Any other value is recorded in d['unserved'] and driver.Simulation refuses the run (Unsupported):
# a run directory must never complete with reference-format files and different physics.
_OFF = ('n', '0', 'f', '.false.', 'false')
UNSERVED = {
    # Hamiltonian terms outside SURVEY 8 (a)
    'do_dip': _OFF, 'pd': None, 'biqdm': None, 'chir': None, 'sa': None, 'ring': None, 'fourx': None,
    'do_lsf': _OFF, 'ind_mom_flag': _OFF, 'mult_axis': _OFF, 'random_anisotropy': _OFF, 'exc_inter': _OFF, 'do_ewald': _OFF,
    'do_macro_cells': _OFF, 'do_efield': _OFF, 'efield': None, 'demag': _OFF, 'do_sparse': _OFF, 'exchangedlm': None,
    'jij_scale': ('1', '1.0', '1.d0', '1.0d0'), 'dm_scale': ('1', '1.0', '1.d0', '1.0d0'), 'ea_model': _OFF, 'rdm_model': _OFF,
    'locfield': _OFF, 'siteatomfield': None, 'do_fixed_mom': _OFF, 'do_mom_legacy': _OFF, 'multiscale': None,
    # dynamics outside the two solvers / the constant uniform temperature
    'stt': _OFF, 'do_she': _OFF, 'do_sot': _OFF, 'do_qhb': _OFF, 'gradtemp': ('0',), 'grad': _OFF,
    'do_3tm': _OFF, 'do_site_damping': _OFF, 'do_site_ip_damping': _OFF, 'damping2': ('0', '0.0', '0.d0', '0.0d0'),
    'ip_damping2': ('0', '0.0', '0.d0', '0.0d0'), 'compensate_drift': ('0',), 'llg': ('1',), 'relaxtime': ('0', '0.0', '0.d0'),
    'do_ld': _OFF, 'do_sld': _OFF, 'do_gneb': _OFF, 'do_kmc': _OFF, 'do_wl': _OFF, 'para_rng': _OFF, 'ziggurat': ('y', 't'),
}
# measurements of the reference this driver does not write: the run is the same physics, the files are absent.  Recorded in
# d['unwritten']; the driver warns once per keyword (the reference's own regression directories switch many of them on).
UNWRITTEN = {
    'do_sc': _OFF, 'do_ams': _OFF, 'do_magdos': _OFF, 'do_autocorr': _OFF, 'do_currents': _OFF, 'do_pol': _OFF, 'do_loc_pol': _OFF,
    'do_prn_beff': _OFF, 'do_prn_binteff': _OFF, 'do_prn_torques': _OFF, 'do_prn_induced': _OFF, 'do_stiffness': _OFF,
    'do_dm_stiffness': _OFF, 'do_larmor_loc': _OFF, 'do_larmor_dos': _OFF, 'do_skyno_den': _OFF, 'do_skyno_cmass': _OFF,
    'do_proj_skyno': _OFF, 'do_mc_avrg': _OFF, 'do_chiral': _OFF, 'do_thermfield': _OFF, 'do_spintemp': _OFF, 'do_bls': _OFF,
    'do_sc_local_axis': _OFF, 'do_sc_proj': _OFF, 'do_sc_projch': _OFF, 'do_sc_bimag': _OFF, 'do_sc_complex': _OFF,
    'do_sc_dosonly': _OFF, 'do_sc_proj_axis': _OFF, 'do_bls_local_axis': _OFF, 'do_qt_traj': _OFF, 'do_connected': _OFF,
    'do_projch_avrg': _OFF,
}
# keywords of the reference that change nothing on this path (printing, memory, tolerances of other modes): accepted silently
_NEUTRAL = {'do_meminfo', 'do_storeham', 'do_hoc_debug', 'evolveout', 'heisout', 'logsamp', 'real_time_measure', 'gpu_rng', 'use_vsl',
            'block_size', 'block_size_x', 'block_size_y', 'block_size_z', 'mseed', 'set_landeg', 'mcavrg_step', 'mcavrg_buff',
            'natoms', 'ntypes', 'do_anisotropy', 'do_prn_poscar', 'prn_ovf', 'read_ovf', 'ip_nstep', 'calc_jtensor',
            # parameters of measurements that are themselves recorded as unwritten / of modes that are refused
            'sc_nstep', 'sc_step', 'sc_sep', 'sc_average', 'sc_window_fun', 'sc_local_axis_mix', 'qpoints', 'qfile', 'bls_nstep',
            'bls_step', 'ene_step', 'ene_buff', 'acfile', 'max_pol_nn', 'jvec', 'adibeta', 'beff_step', 'beff_buff', 'binteff_step',
            'binteff_buff', 'torques_step', 'torques_buff', 'thermfield_step', 'thermfield_buff', 'larm_step', 'larm_buff',
            'larm_dos_size', 'pol_step', 'pol_buff', 'current_step', 'current_buff', 'ind_step', 'ind_buff', 'spintemp_step',
            'magdos_freq', 'magdos_hfreq', 'magdos_lfreq', 'magdos_sigma', 'eta_max', 'eta_min'}


def read_inpsd(path):
    """Keywords of the hot-path slice.  A keyword of the reference that switches on physics or a measurement outside this slice
    is recorded in d['unserved'] (driver.Simulation raises Unsupported for it), a measurement this driver does not write in
    d['unwritten'], every other unrecognised keyword in d['ignored'] (the driver warns about both kinds)."""
    d = defaults()
    d['unserved'], d['unwritten'], d['ignored'] = [], [], []
    base = os.path.dirname(os.path.abspath(path))
    with open(path) as fh:
        lines = fh.read().splitlines()
    i = 0

    def next_row():
        nonlocal i
        while i < len(lines):
            t = lines[i].split()
            i += 1
            if t:
                return t
        raise InputError('unexpected end of %s' % path)

    while i < len(lines):
        t = lines[i].split()
        i += 1
        if not t or t[0].startswith(('#', '!', '%', '*')):
            continue
        key, v = t[0].lower(), t[1:]
        try:
            if key == 'simid':
                d['simid'] = v[0][:8]
            elif key == 'ncell':
                d['ncell'] = tuple(int(_num(x)) for x in v[:3])
            elif key == 'bc':
                d['bc'] = tuple(x.upper()[0] for x in v[:3])
            elif key == 'cell':
                rows = [v[:3]]
                while len(rows) < 3:
                    rows.append(next_row()[:3])
                d['cell'] = np.array([[_num(x) for x in r] for r in rows])
            elif key in _FILES:
                d[key] = os.path.normpath(os.path.join(base, v[0]))
            elif key in _SCALAR_INT:
                d[key] = int(_num(v[0]))
            elif key in _SCALAR_REAL:
                d[key] = _num(v[0])
            elif key in _SCALAR_FLAG:
                d[key] = _flag(v[0])
            elif key in ('hfield', 'ip_hfield'):
                d[key] = tuple(_num(x) for x in v[:3])
            elif key == 'ip_nphase':          # rows: nstep  Temp  timestep  damping   (inputhandler.f90:788-850)
                n = int(_num(v[0]))
                d['ip_nphase'] = []
                for _ in range(n):
                    r = next_row()
                    d['ip_nphase'].append((int(_num(r[0])), _num(r[1]), _num(r[2]), _num(r[3])))
            elif key == 'ip_mcanneal':        # rows: nsweeps  Temp                      (inputhandler.f90:887-915)
                n = int(_num(v[0]))
                d['ip_mcanneal'] = []
                for _ in range(n):
                    r = next_row()
                    d['ip_mcanneal'].append((int(_num(r[0])), _num(r[1])))
            elif key == 'ntraj':              # rows: atom  step  buffer                 (prn_trajectories.f90:476-495)
                n = int(_num(v[0]))
                d['trajectories'] = []
                for _ in range(n):
                    r = next_row()
                    d['trajectories'].append((int(_num(r[0])), int(_num(r[1])), int(_num(r[2]))))
            elif key == 'map_multiple':
                d['map_multiple'] = _flag(v[0]) in ('T', 'Y')
            elif key == 'do_cumu' or key == 'skyno':
                pass
            elif key in UNSERVED:
                off = UNSERVED[key]
                val = v[0].lower().strip('.') if v else ''
                if off is None or val not in [o.strip('.') for o in off]:
                    d['unserved'].append((key, ' '.join(v)))
            elif key in UNWRITTEN:
                val = v[0].lower().strip('.') if v else ''
                if val not in [o.strip('.') for o in UNWRITTEN[key]]:
                    d['unwritten'].append(key)
            elif key not in _NEUTRAL:
                d['ignored'].append(key)
        except (IndexError, ValueError) as exc:
            raise InputError('cannot read keyword %s in %s: %s' % (key, path, exc))
    if d['do_bpulse'] not in (0, 1, 2, 3, 4):
        d['unserved'].append(('do_bpulse', str(d['do_bpulse'])))     # 5 / 6: site-dependent static fields from files
    if d['do_cumu'] not in ('Y', 'N'):
        d['unwritten'].append('do_cumu ' + d['do_cumu'])      # 'A' (cumulants of every ensemble) is not written here
    if d['skyno'] not in ('N', 'T'):
        d['unwritten'].append('skyno ' + d['skyno'])          # 'Y' (finite-difference skyrmion number): only the triangulated 'T' form
    if d['ipsdealgh'] == -1:
        d['ipsdealgh'] = d['sdealgh']             # uppasd.f90:885
    return d


# ---------------------------------------------------------------------------------------------------------------
# structure files
# ---------------------------------------------------------------------------------------------------------------
def read_posfile(path, cell, posfiletype='C'):
    """bas(3,NA) in Cartesian units of the lattice constant, atype(NA)."""
    rows = _data_rows(path)
    na = max(int(r[0]) for r in rows)
    bas = np.zeros((3, na))
    atype = np.zeros(na, dtype=np.int32)
    for r in rows:
        i = int(r[0]) - 1
        p = np.array([_num(x) for x in r[2:5]])
        bas[:, i] = p[0] * cell[0] + p[1] * cell[1] + p[2] * cell[2] if posfiletype == 'D' else p
        atype[i] = int(r[1])
    return bas, atype


def read_momfile(path, na, landeg_glob=2.0):
    """ammom(NA), aemom(3,NA) normalised like read_moments does, Landeg(NA)."""
    ammom = np.zeros(na)
    aemom = np.zeros((3, na))
    for r in _data_rows(path):
        i = int(r[0]) - 1
        ammom[i] = _num(r[2])
        e = np.array([_num(x) for x in r[3:6]])
        aemom[:, i] = e / np.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
    return ammom, aemom, np.full(na, landeg_glob)


def read_pairfile(path, atype, bas, cell, maptype, posfiletype, ncomp):
    """Shell table of a pair interaction file: nn(NT), redcoord(NT,maxshell,3), xc(ncomp,NT,maxshell),
    nntype(NT,maxshell).  A later line that repeats a shell vector (within 1e-5 squared) overwrites its value."""
    nt = int(atype.max())
    shells = [[] for _ in range(nt)]
    for r in _data_rows(path):
        isite, jsite = int(r[0]), int(r[1])
        rt = [_num(x) for x in r[2:5]]
        val = [_num(x) for x in r[5:5 + ncomp]]
        if maptype == 2:      # bgfm style: vector between the two sites + cell translation (getNeighVec)
            vec = np.array([bas[a, jsite - 1] - bas[a, isite - 1] + cell[0][a] * rt[0] + cell[1][a] * rt[1] + cell[2][a] * rt[2]
                            for a in range(3)])
        elif posfiletype == 'D':
            vec = np.array([rt[0] * cell[0][a] + rt[1] * cell[1][a] + rt[2] * cell[2][a] for a in range(3)])
        else:
            vec = np.array(rt)
        lst = shells[int(atype[isite - 1]) - 1]
        for ent in lst:
            if ((vec - ent[0]) ** 2).sum() < 1.0e-5:
                ent[1] = val
                break
        else:
            lst.append([vec, val, int(atype[jsite - 1])])
    nn = np.array([len(s) for s in shells], dtype=np.int32)
    ms = max(1, int(nn.max()))
    red = np.zeros((nt, ms, 3))
    xc = np.zeros((ncomp, nt, ms))
    nntype = np.zeros((nt, ms), dtype=np.int32)
    for t, lst in enumerate(shells):
        for s, (vec, val, jt) in enumerate(lst):
            red[t, s] = vec
            xc[:, t, s] = val
            nntype[t, s] = jt
    return nn, red, xc, nntype


# ---- random alloys (do_ralloy 1): the three files carry a chemical-type column (inputhandler_ext.f90:139-216, 228-321, 487-531)
def read_posfile_alloy(path, cell, posfiletype='C'):
    """rows `site type chem concentration x y z`: bas(3,NA), atype(NA), nch(NA), chconc(NA,Nchmax)"""
    rows = _data_rows(path)
    na = max(int(r[0]) for r in rows)
    nchmax = max(int(r[2]) for r in rows)
    bas, atype = np.zeros((3, na)), np.zeros(na, dtype=np.int32)
    nch, chconc = np.zeros(na, dtype=np.int32), np.zeros((na, nchmax))
    for r in rows:
        i, ich = int(r[0]) - 1, int(r[2])
        p = np.array([_num(x) for x in r[4:7]])
        bas[:, i] = p[0] * cell[0] + p[1] * cell[1] + p[2] * cell[2] if posfiletype == 'D' else p
        atype[i] = int(r[1])
        nch[i] = max(nch[i], ich)
        chconc[i, ich - 1] = _num(r[3])
    return bas, atype, nch, chconc


def read_momfile_alloy(path, na, nchmax, landeg_glob=2.0):
    """rows `site chem moment ex ey ez`: ammom(NA,Nchmax), aemom(3,NA,Nchmax) normalised, Landeg(NA,Nchmax)"""
    ammom, aemom = np.zeros((na, nchmax)), np.zeros((3, na, nchmax))
    for r in _data_rows(path):
        i, c = int(r[0]) - 1, int(r[1]) - 1
        ammom[i, c] = _num(r[2])
        e = np.array([_num(x) for x in r[3:6]])
        aemom[:, i, c] = e / np.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
    return ammom, aemom, np.full((na, nchmax), landeg_glob)


def read_pairfile_alloy(path, atype, nchmax, bas, cell, maptype, posfiletype):
    """rows `isite jsite ichem jchem r1 r2 r3 J`: nn(NT), redcoord(NT,maxshell,3), xc(NT,maxshell,Nchmax,Nchmax), nntype(NT,maxshell);
    shells are told apart by their vector exactly as in read_pairfile, the value is filed under (ichem, jchem)"""
    nt = int(atype.max())
    shells = [[] for _ in range(nt)]
    for r in _data_rows(path):
        isite, jsite, ichem, jchem = int(r[0]), int(r[1]), int(r[2]), int(r[3])
        rt = [_num(x) for x in r[4:7]]
        if maptype == 2:
            vec = np.array([bas[a, jsite - 1] - bas[a, isite - 1] + cell[0][a] * rt[0] + cell[1][a] * rt[1] + cell[2][a] * rt[2]
                            for a in range(3)])
        elif posfiletype == 'D':
            vec = np.array([rt[0] * cell[0][a] + rt[1] * cell[1][a] + rt[2] * cell[2][a] for a in range(3)])
        else:
            vec = np.array(rt)
        lst = shells[int(atype[isite - 1]) - 1]
        for ent in lst:
            if ((vec - ent[0]) ** 2).sum() < 1.0e-5:
                ent[1][ichem - 1, jchem - 1] = _num(r[7])
                break
        else:
            m = np.zeros((nchmax, nchmax))
            m[ichem - 1, jchem - 1] = _num(r[7])
            lst.append([vec, m, int(atype[jsite - 1])])
    nn = np.array([len(s) for s in shells], dtype=np.int32)
    ms = max(1, int(nn.max()))
    red, xc = np.zeros((nt, ms, 3)), np.zeros((nt, ms, nchmax, nchmax))
    nntype = np.zeros((nt, ms), dtype=np.int32)
    for t, lst in enumerate(shells):
        for s, (vec, m, jt) in enumerate(lst):
            red[t, s], xc[t, s], nntype[t, s] = vec, m, jt
    return nn, red, xc, nntype


def read_tensorfile(path, atype, bas, cell, maptype, posfiletype):
    """Exchange file in tensor format (do_jtensor 1; read_exchange_tensor_base, inputhandler_ext.f90:658-740): nine numbers
    per line are read into j_tmp(3,3) in Fortran (column-major) order and then transposed, i.e. the file lists the tensor
    row by row; the neighbour type is not distinguished.  Returns nn, redcoord, xc(9, NT, maxshell) with J(a,b) at component
    a + 3 b (the Fortran storage order of j_tens(3,3,...)), and None for nntype."""
    nn, red, xc, _ = read_pairfile(path, atype, bas, cell, maptype, posfiletype, 9)
    out = np.zeros_like(xc)
    for a in range(3):
        for b in range(3):
            out[a + 3 * b] = xc[3 * a + b]
    return nn, red, out, None


def read_kfile(path, na):
    """anisotropytype(NA), anisotropy(NA,6) = K1, K2, ex, ey, ez, ratio."""
    atyp = np.zeros(na, dtype=np.int32)
    an = np.zeros((na, 6))
    for r in _data_rows(path)[:na]:
        i = int(r[0]) - 1
        atyp[i] = int(r[1])
        an[i] = [_num(x) for x in r[2:8]]
    return atyp, an


def read_restart(path, natom, mensemble):
    """restart.<simid>.out -> rstep, emom(3,N,M), mmom(N,M)."""
    emom = np.zeros((3, natom, mensemble), order='F')
    mmom = np.zeros((natom, mensemble), order='F')
    rstep = 0
    for r in _data_rows(path):
        it, k, i = int(r[0]), int(r[1]) - 1, int(r[2]) - 1
        rstep = it
        mmom[i, k] = _num(r[3])
        e = np.array([_num(x) for x in r[4:7]])
        # read_mag_conf_std normalises every row (restart.f90:354-357, f_normalize_vec = vec / norm2(vec)): the file holds 9 digits
        emom[:, i, k] = e / np.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
    return rstep, emom, mmom


# ---------------------------------------------------------------------------------------------------------------
# Fortran edit descriptors
# ---------------------------------------------------------------------------------------------------------------
def es16_8(x):
    """ES16.8 like gfortran prints it (three-digit exponents drop the E)."""
    s = '%16.8E' % x
    mant, exp = s.split('E')
    e = int(exp)
    if abs(e) > 99:
        return ('%s%+04d' % (mant, e)).rjust(16)
    return s


def _row(first, vals):
    return first + ''.join(es16_8(v) for v in vals) + '\n'


class OutputFiles:
    """The measurement files of one run directory (append mode like the reference; truncated when the run starts)."""

    def __init__(self, directory, simid):
        self.dir, self.simid = directory, simid.strip()
        self._started = set()

    def _open(self, name):
        path = os.path.join(self.dir, name)
        mode = 'a' if name in self._started else 'w'
        self._started.add(name)
        return open(path, mode)

    def averages(self, rows):
        """rows: (iter, mx, my, mz, m, mstdv) -- prn_avrg, format 10004 = (i8,6es16.8)"""
        name = 'averages.%s.out' % self.simid
        new = name not in self._started
        with self._open(name) as fh:
            if new:
                fh.write('%8s%16s%16s%16s%16s%16s\n' % ('#Iter', '<M>_x', '<M>_y', '<M>_z', '<M>', 'M_{stdv}'))
            for r in rows:
                fh.write(_row('%8d' % r[0], r[1:]))

    def cumulants(self, row):
        """(count, <M>, <M^2>, <M^4>, U, chi, Cv, <E>, <E_exc>, <E_lsf>) -- calc_and_print_cumulant, (i8,10es16.8)"""
        name = 'cumulants.%s.out' % self.simid
        new = name not in self._started
        with self._open(name) as fh:
            if new:
                fh.write('%8s%16s%16s%16s%16s%16s%16s%16s%16s%16s\n' % ('#Iter', '<M>', '<M^2>', '<M^4>', 'U_{Binder}', '\\chi',
                                                                        'C_v(tot)', '<E>', '<E_{exc}>', '<E_{lsf}>'))
            fh.write(_row('%8d' % row[0], row[1:]))

    def totenergy(self, it, terms):
        """terms: dict with tot, exc, ani, dm, bq, ext (mRy per atom, ensemble means) -- energy.f90:404-430, (i8,13es16.8)"""
        name = 'totenergy.%s.out' % self.simid
        new = name not in self._started
        with self._open(name) as fh:
            if new:
                fh.write('%8s' % '#Iter' + ''.join('%16s' % h for h in ('Tot', 'Exc', 'Ani', 'DM', 'PD', 'BiqDM', 'BQ', 'Dip',
                                                                           'Zeeman', 'LSF', 'Chir', 'Ring', 'SA')) + '\n')
            z = 0.0
            fh.write(_row('%8d' % it, [terms['tot'], terms['exc'], terms['ani'], terms['dm'], z, z, terms['bq'], z,
                                       terms['ext'], z, z, z, z]))

    def projavgs(self, rows):
        """rows: (iter, proj, <M>, M_stdv, <M>_x, <M>_y, <M>_z) -- prn_proj_avrg, format 10004 = (i8,i8,5es16.8)"""
        name = 'projavgs.%s.out' % self.simid
        new = name not in self._started
        with self._open(name) as fh:
            if new:
                fh.write('%8s%s%16s%16s%16s%16s%16s\n' % ('#Iter', 'Proj', '<M>', 'M_{stdv}', '<M>_x', '<M>_y', '<M>_z'))
            for r in rows:
                fh.write(_row('%8d%8d' % (r[0], r[1]), r[2:]))

    def projcumulants(self, rows):
        """rows: (count, type, <M>, <M^2>, <M^4>, U, chi) -- calc_and_print_cumulant_proj, format 10004 = (i8,i6,5es16.8)"""
        name = 'projcumulants.%s.out' % self.simid
        new = name not in self._started
        with self._open(name) as fh:
            if new:
                fh.write('%8s%6s%16s%16s%16s%16s%16s\n' % ('# ter.', 'Type', '<M>', '<M^2>', '<M^4>', 'U_{Binder}', '\\chi'))
            for r in rows:
                fh.write(_row('%8d%6d' % (r[0], r[1]), r[2:]))

    def sknumber(self, rows):
        """rows: (iter, Skx num, Skx avg, Skx std) -- prn_skyno, format 240 = (i8,2x,5f16.8)"""
        name = 'sknumber.%s.out' % self.simid
        new = name not in self._started
        with self._open(name) as fh:
            if new:
                fh.write('%8s  %10s  %10s  %10s\n' % ('# Iter', 'Skx num', 'Skx avg', 'Skx std'))
            for r in rows:
                fh.write('%8d  ' % r[0] + ''.join('%16.8f' % v for v in r[1:]) + '\n')

    def trajectory(self, atom, ens, rows):
        """rows: (iter, ex, ey, ez, m) -- prn_traj, format 10002 = (i8,2x,i8,2x,2x,4es16.8)"""
        name = 'trajectory.%s.%03d.%1d.out' % (self.simid, atom, ens)
        with self._open(name) as fh:
            for r in rows:
                fh.write('%8d  %8d    ' % (r[0], atom) + ''.join(es16_8(v) for v in r[1:]) + '\n')

    def restart(self, mstep, mode, emom, mmom):
        """restart.<simid>.out -- prn_mag_conf_iter type 'R' (restart.f90:186-246), rewritten every time"""
        n, m = mmom.shape
        with open(os.path.join(self.dir, 'restart.%s.out' % self.simid), 'w') as fh:
            fh.write('#' * 80 + '\n')
            fh.write('# File type: R\n# Simulation type: %s\n' % mode)
            fh.write('# Number of atoms:  %8d\n# Number of ensembles:  %8d\n' % (n, m))
            fh.write('#' * 80 + '\n')
            fh.write('%8s%8s%8s%16s%16s%16s%16s\n' % ('# iter', 'ens', 'iatom', '|Mom|', 'M_x', 'M_y', 'M_z'))
            for k in range(m):
                for i in range(n):
                    fh.write('%8d%8d%8d  ' % (mstep, k + 1, i + 1) + es16_8(mmom[i, k]) + es16_8(emom[0, i, k])
                             + es16_8(emom[1, i, k]) + es16_8(emom[2, i, k]) + '\n')

    def coord(self, coord, atype, anumb):
        """coord.<simid>.out -- printhamiltonian / geometry output: index, x, y, z, type, number in cell"""
        with open(os.path.join(self.dir, 'coord.%s.out' % self.simid), 'w') as fh:
            for i in range(coord.shape[1]):
                fh.write('%7d%12.6f%12.6f%12.6f%6d%6d\n' % (i + 1, coord[0, i], coord[1, i], coord[2, i], atype[i], anumb[i]))


def read_out(path):
    """rows of a .out file as lists of floats, comment lines skipped (what tests/extractoutput.py does)"""
    out = []
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t or t[0].startswith('#'):
                continue
            out.append([float(x) for x in t])
    return out
