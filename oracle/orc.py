"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/liborc.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product (uppasd_b200/) never does.

`build_system()` chains the restated setup routines in the order the reference's `setup_simulation` does
(source/uppasd.f90:914-1200): geometry -> read_exchange/dm/bq/anisotropy -> setup_hamiltonian ->
setup_moment -> magninit.  `sd_run()` replays `sd_mphase` (source/sd_driver.f90:517-849): measure, then
field / evolve_first / field / evolve_second / moment_update.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# source/Parameters/constants.f90:14-29
CONST = dict(gama=1.760859644e11, k_bolt=1.38064852e-23, mub=9.274009994e-24, mry=2.179872325e-21)
# aunits Y (uppasd.f90:794-797 -> change_constants, inputhandler.f90:1685-1704): model Hamiltonians in units where every
# physical constant on this path is 1
AUNITS = dict(gama=1.0, k_bolt=1.0, mub=1.0, mry=1.0)


def consts(S):
    """the constants a system was mounted with (S['const']; SI unless the input says aunits Y)"""
    return S.get('const', CONST)


def build(force=False):
    so = os.path.join(_HERE, 'liborc.so')
    srcs = [os.path.join(_HERE, f) for f in ('orc_setup.cpp', 'orc_dynamics.cpp', 'orc_rng.cpp')]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_nm_create.restype = C.c_void_p
        _LIB.orc_effective_field.restype = C.c_double
        _LIB.orc_rng_raw32.restype = C.c_uint
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _d(x):
    return C.c_double(x)


class OrcHam(C.Structure):
    _fields_ = [('Natom', C.c_int), ('Mensemble', C.c_int), ('nHam', C.c_int), ('max_no_neigh', C.c_int),
                ('nlist', C.c_void_p), ('nlistsize', C.c_void_p), ('ncoup', C.c_void_p), ('aHam', C.c_void_p),
                ('do_dm', C.c_int), ('max_no_dmneigh', C.c_int), ('dmlist', C.c_void_p),
                ('dmlistsize', C.c_void_p), ('dm_vect', C.c_void_p),
                ('do_bq', C.c_int), ('nn_bq_tot', C.c_int), ('bqlist', C.c_void_p), ('bqlistsize', C.c_void_p),
                ('j_bq', C.c_void_p),
                ('do_anisotropy', C.c_int), ('taniso', C.c_void_p), ('eaniso', C.c_void_p),
                ('kaniso', C.c_void_p), ('sb', C.c_void_p), ('do_jtensor', C.c_int)]


def neighbour_table(S, nn, redcoord, xc, nntype, sym, hdim, lexp, do_sortcoup=False, map_multiple=False):
    """setup_nm + setup_neighbour_hamiltonian for one pair interaction.  Returns dict(list, listsize, coup, z)."""
    L = lib()
    N, NT, NA = S['Natom'], S['NT'], S['NA']
    N1, N2, N3 = S['ncell']
    cell = S['cell']
    c1, c2, c3 = (np.ascontiguousarray(cell[i], dtype=np.float64) for i in range(3))
    nn = np.ascontiguousarray(nn, dtype=np.int32)
    ms = redcoord.shape[1]
    redcoord = np.asfortranarray(redcoord, dtype=np.float64)
    nnt = np.asfortranarray(nntype, dtype=np.int32) if nntype is not None else None
    h = L.orc_nm_create(N, NT, NA, N1, N2, N3, _p(c1), _p(c2), _p(c3), C.c_char(S['bc'][0].encode()),
                        C.c_char(S['bc'][1].encode()), C.c_char(S['bc'][2].encode()), _p(S['atype']), _p(S['bas']),
                        ms, sym, _p(nn), _p(redcoord), _p(nnt))
    h = C.c_void_p(h)
    z = L.orc_nm_max_no_neigh(h)
    NH = S['nHam']
    nlist = np.zeros((z, N), dtype=np.int32, order='F')
    nlistsize = np.zeros(NH, dtype=np.int32)
    ncoup = np.zeros((hdim, z, NH), order='F')
    xc = np.asfortranarray(xc, dtype=np.float64)
    L.orc_mount(h, N, NT, NA, NH, _p(S['anumb']), _p(S['atype']), z, _p(nn), _p(xc), _p(S['ammom_inp']), hdim, lexp,
                int(do_sortcoup), int(map_multiple), _d(consts(S)['mry']), _d(consts(S)['mub']), _p(nlistsize), _p(nlist),
                _p(ncoup))
    me, mnn = L.orc_nm_max_no_equiv(h), L.orc_nm_maxnn(h)
    nm_cell = np.zeros((me, ms, NA), dtype=np.int32, order='F')
    nm_trunk = np.zeros((3, me, ms, NA), dtype=np.int32, order='F')
    nnm_cell = np.zeros((ms, NA), dtype=np.int32, order='F')
    L.orc_nm_get_stencil(h, _p(nm_cell), _p(nm_trunk), _p(nnm_cell))
    L.orc_nm_free(h)
    if hdim == 1:
        ncoup = np.asfortranarray(ncoup[0])
    return dict(list=nlist, listsize=nlistsize, coup=ncoup, z=z, nm_cell=nm_cell, nm_trunk=nm_trunk,
                nnm_cell=nnm_cell, maxnn=mnn)


def build_system(inp, bas, atype_inp, ammom_inp, aemom_inp, landeg_ch, exchange, dm=None, bq=None, aniso=None):
    """exchange/dm/bq: callables (S) -> (nn, redcoord, xc, nntype) evaluated AFTER the basis has been folded
    (read_exchange runs after setup_geometry, source/uppasd.f90:921-930), or ready tuples."""
    L = lib()
    N1, N2, N3 = inp['ncell']
    NA = bas.shape[1]
    N = NA * N1 * N2 * N3
    M = inp['mensemble']
    cell = np.array(inp['cell'], dtype=np.float64)
    S = dict(Natom=N, NA=NA, NT=int(atype_inp.max()), ncell=(N1, N2, N3), cell=cell, bc=inp['bc'], Mensemble=M)
    S['const'] = dict(AUNITS if inp.get('aunits', 'N') == 'Y' else CONST)
    S['bas'] = np.asfortranarray(bas, dtype=np.float64).copy(order='F')
    S['atype_inp'] = np.ascontiguousarray(atype_inp, dtype=np.int32)
    anumb_inp = np.arange(1, NA + 1, dtype=np.int32)
    S['coord'] = np.zeros((3, N), order='F')
    S['atype'] = np.zeros(N, dtype=np.int32)
    S['anumb'] = np.zeros(N, dtype=np.int32)
    c1, c2, c3 = (np.ascontiguousarray(cell[i]) for i in range(3))
    L.orc_setup_geometry(NA, N1, N2, N3, _p(c1), _p(c2), _p(c3), _p(S['bas']), _p(S['atype_inp']), _p(anumb_inp),
                         _p(S['coord']), _p(S['atype']), _p(S['anumb']))
    S['ammom_inp'] = np.ascontiguousarray(ammom_inp, dtype=np.float64)
    reduced = inp['do_reduced'] == 'Y'
    S['nHam'] = NA if reduced else N
    S['aHam'] = (S['anumb'].copy() if reduced else np.arange(1, N + 1, dtype=np.int32))
    sortc = inp['do_sortcoup'] == 'Y'
    ex = exchange(S) if callable(exchange) else exchange
    S['do_jtensor'] = int(inp.get('do_jtensor', 0))
    if S['do_jtensor'] == 1:
        # tensorial exchange: same neighbour map, nine couplings per pair, lexp = 1 (hamiltonianinit.f90:412-432)
        S['exchange'] = neighbour_table(S, ex[0], ex[1], ex[2], None, inp['sym'], 9, 1, sortc, inp['map_multiple'])
    else:
        S['exchange'] = neighbour_table(S, ex[0], ex[1], ex[2], ex[3], inp['sym'], 1, 1, sortc, inp['map_multiple'])
    S['dm'] = None
    if dm is not None:
        t = dm(S) if callable(dm) else dm
        S['dm'] = neighbour_table(S, t[0], t[1], t[2], None, 0, 3, 1, sortc, inp['map_multiple'])
    S['bq'] = None
    if bq is not None:
        t = bq(S) if callable(bq) else bq
        S['bq'] = neighbour_table(S, t[0], t[1], t[2], None, inp['sym'], 1, 2, sortc, inp['map_multiple'])
    S['aniso'] = None
    if aniso is not None:
        atyp, an = aniso
        ta = np.zeros(N, dtype=np.int32)
        ea = np.zeros((3, N), order='F')
        ka = np.zeros((2, N), order='F')
        sb = np.zeros(N)
        L.orc_setup_anisotropies(N, NA, _p(S['anumb']), _p(np.ascontiguousarray(atyp, dtype=np.int32)),
                                 _p(np.asfortranarray(an)), _p(S['ammom_inp']), _d(consts(S)['mry']), _d(consts(S)['mub']),
                                 _p(ta), _p(ea), _p(ka), _p(sb))
        S['aniso'] = dict(taniso=ta, eaniso=ea, kaniso=ka, sb=sb)
    for k in ('mmom', 'mmom0', 'mmomi'):
        S[k] = np.zeros((N, M), order='F')
    S['Landeg'] = np.zeros(N)
    S['emom'] = np.zeros((3, N, M), order='F')
    S['emomM'] = np.zeros((3, N, M), order='F')
    L.orc_setup_moments(N, M, NA, N1, N2, N3, _p(S['ammom_inp']), _p(np.asfortranarray(aemom_inp)),
                        _p(np.ascontiguousarray(landeg_ch, dtype=np.float64)), _p(S['mmom']), _p(S['mmom0']),
                        _p(S['mmomi']), _p(S['Landeg']), _p(S['emom']), _p(S['emomM']))
    S['external_field'] = np.zeros((3, N, M), order='F')
    for a in range(3):
        S['external_field'][a, :, :] = inp['hfield'][a]
    return S


def setup_chemicaldata(NA, ncell, nch, chconc, tseed):
    """setup_chemicaldata (geometry.f90:190-329), do_ralloy 1: chemical type of every site of the full supercell.
    The generator is re-initialised with tseed just before (uppasd.f90:903-906); per basis site Ncell uniforms are drawn and
    the cells are dealt to the species in the order of repeated maxloc(rn) with the maximum zeroed after each pick (:266-270),
    i.e. by descending random number, the lowest index first among equal values; species ich gets qch = nint(conc * Ncell)
    consecutive picks (:251-265,271-276).  Returns achtype(Natom_full), 0 = vacancy (dilute system)."""
    n1, n2, n3 = ncell
    ncellt = n1 * n2 * n3
    achtype = np.zeros(NA * ncellt, dtype=np.int32)
    rng_init(tseed)
    for ia in range(1, NA + 1):
        qch = [int(np.rint(chconc[ia - 1, ich] * ncellt)) if ich < nch[ia - 1] else 0 for ich in range(chconc.shape[1])]
        rn = rng_uniform(ncellt)
        atoms = np.argsort(-rn, kind='stable') + 1            # atoms(i, ia) = i-th largest random number's cell
        ns = 1
        for ich in range(1, nch[ia - 1] + 1):
            ne = ns + qch[ich - 1] - 1
            for i in range(ns, min(ne, ncellt) + 1):
                achtype[(atoms[i - 1] - 1) * NA + ia - 1] = ich
            ns = ne + 1
    return achtype


def build_alloy_system(inp, bas, atype_inp, nch, chconc, ammom_inp, aemom_inp, landeg_ch, exchange):
    """build_system for a random alloy (do_ralloy 1, non-dilute, scalar exchange only): occupancy (setup_chemicaldata), the
    neighbour map of the full supercell, the mount with chemistry-dependent couplings (orc_mount_alloy) and the moments of
    setup_moment / magninit's Initmag 3 for alloys (magnetizationinit.f90:244-258,586-597).  nHam = Natom."""
    L = lib()
    N1, N2, N3 = inp['ncell']
    NA = bas.shape[1]
    N = NA * N1 * N2 * N3
    M = inp['mensemble']
    cell = np.array(inp['cell'], dtype=np.float64)
    S = dict(Natom=N, NA=NA, NT=int(atype_inp.max()), ncell=(N1, N2, N3), cell=cell, bc=inp['bc'], Mensemble=M)
    S['const'] = dict(AUNITS if inp.get('aunits', 'N') == 'Y' else CONST)
    S['bas'] = np.asfortranarray(bas, dtype=np.float64).copy(order='F')
    S['atype_inp'] = np.ascontiguousarray(atype_inp, dtype=np.int32)
    anumb_inp = np.arange(1, NA + 1, dtype=np.int32)
    S['coord'] = np.zeros((3, N), order='F')
    S['atype'] = np.zeros(N, dtype=np.int32)
    S['anumb'] = np.zeros(N, dtype=np.int32)
    c1, c2, c3 = (np.ascontiguousarray(cell[i]) for i in range(3))
    L.orc_setup_geometry(NA, N1, N2, N3, _p(c1), _p(c2), _p(c3), _p(S['bas']), _p(S['atype_inp']), _p(anumb_inp),
                         _p(S['coord']), _p(S['atype']), _p(S['anumb']))
    achtype = setup_chemicaldata(NA, (N1, N2, N3), nch, chconc, inp['tseed'])
    if (achtype == 0).any():
        raise ValueError('dilute alloys (vacant sites) are not restated')
    # geometry.f90:285-300 for a fully occupied supercell: acellnumb = identity
    S['achtype'] = achtype
    S['asite_ch'] = S['anumb'].copy()
    S['atype_ch'] = S['atype'].copy()
    S['achem_ch'] = achtype.copy()
    S['nHam'] = N
    S['aHam'] = np.arange(1, N + 1, dtype=np.int32)
    S['do_jtensor'] = 0
    nn, redcoord, xc, nntype = exchange(S) if callable(exchange) else exchange
    nn = np.ascontiguousarray(nn, dtype=np.int32)
    ms = redcoord.shape[1]
    redcoord = np.asfortranarray(redcoord, dtype=np.float64)
    nnt = np.asfortranarray(nntype, dtype=np.int32) if nntype is not None else None
    h = C.c_void_p(L.orc_nm_create(N, S['NT'], NA, N1, N2, N3, _p(c1), _p(c2), _p(c3), C.c_char(S['bc'][0].encode()),
                                   C.c_char(S['bc'][1].encode()), C.c_char(S['bc'][2].encode()), _p(S['atype']), _p(S['bas']),
                                   ms, inp['sym'], _p(nn), _p(redcoord), _p(nnt)))
    z = L.orc_nm_max_no_neigh(h)
    nlist = np.zeros((z, N), dtype=np.int32, order='F')
    nlistsize = np.zeros(N, dtype=np.int32)
    ncoup = np.zeros((1, z, N), order='F')
    xc = np.asfortranarray(xc, dtype=np.float64)
    am = np.asfortranarray(ammom_inp, dtype=np.float64)
    nchmax = am.shape[1]
    L.orc_mount_alloy(h, N, S['NT'], NA, nchmax, _p(S['atype_ch']), _p(S['asite_ch']), _p(S['achem_ch']), z, _p(nn), _p(xc), _p(am),
                      1, 1, int(inp['do_sortcoup'] == 'Y'), int(inp['map_multiple']), _d(consts(S)['mry']), _d(consts(S)['mub']),
                      _p(nlistsize), _p(nlist), _p(ncoup))
    L.orc_nm_free(h)
    S['exchange'] = dict(list=nlist, listsize=nlistsize, coup=np.asfortranarray(ncoup[0]), z=z)
    S['dm'] = S['bq'] = S['aniso'] = None
    S['ammom_inp'] = am
    # setup_moment (magnetizationinit.f90:586-597) and Initmag 3 (:244-258) for alloys
    site, chem = S['anumb'] - 1, achtype - 1
    mm = np.abs(am[site, chem])
    S['mmom'] = np.asfortranarray(np.repeat(mm[:, None], M, axis=1))
    S['mmom0'] = S['mmom'].copy(order='F')
    S['mmomi'] = np.asfortranarray(1.0 / S['mmom'])
    S['Landeg'] = np.asfortranarray(landeg_ch)[site, chem] * 0.5
    ae = np.asfortranarray(aemom_inp)[:, site, chem]
    S['emom'] = np.asfortranarray(np.repeat(ae[:, :, None], M, axis=2))
    S['emomM'] = np.asfortranarray(S['emom'] * S['mmom'][None])
    S['external_field'] = np.zeros((3, N, M), order='F')
    for a in range(3):
        S['external_field'][a, :, :] = inp['hfield'][a]
    return S


def ham_struct(S):
    """OrcHam view over a system dict (keeps references alive in S['_keep'])."""
    H = OrcHam()
    H.Natom, H.Mensemble, H.nHam = S['Natom'], S['Mensemble'], S['nHam']
    ex = S['exchange']
    H.max_no_neigh = ex['z']
    H.nlist, H.nlistsize, H.ncoup, H.aHam = _p(ex['list']), _p(ex['listsize']), _p(ex['coup']), _p(S['aHam'])
    H.do_jtensor = int(S.get('do_jtensor', 0))          # ncoup then holds j_tens(3,3,z,nHam)
    if S.get('dm') is not None:
        t = S['dm']
        H.do_dm, H.max_no_dmneigh = 1, t['z']
        H.dmlist, H.dmlistsize, H.dm_vect = _p(t['list']), _p(t['listsize']), _p(t['coup'])
    if S.get('bq') is not None:
        t = S['bq']
        H.do_bq, H.nn_bq_tot = 1, t['z']
        H.bqlist, H.bqlistsize, H.j_bq = _p(t['list']), _p(t['listsize']), _p(t['coup'])
    if S.get('aniso') is not None:
        t = S['aniso']
        H.do_anisotropy = 1
        H.taniso, H.eaniso, H.kaniso, H.sb = _p(t['taniso']), _p(t['eaniso']), _p(t['kaniso']), _p(t['sb'])
    return H


def effective_field(S, emomM=None, want_parts=False):
    L = lib()
    H = ham_struct(S)
    N, M = S['Natom'], S['Mensemble']
    emomM = S['emomM'] if emomM is None else np.asfortranarray(emomM)
    beff = np.zeros((3, N, M), order='F')
    b1 = np.zeros((3, N, M), order='F') if want_parts else None
    b2 = np.zeros((3, N, M), order='F') if want_parts else None
    e = L.orc_effective_field(C.byref(H), _p(emomM), _p(S['external_field']), _p(beff), _p(b1), _p(b2),
                              _d(consts(S)['mub']), _d(consts(S)['mry']))
    return (beff, b1, b2, e) if want_parts else (beff, e)


def energy_terms(S, emomM=None):
    """calc_energy (energy.f90:181-398): terms(5, M) = exchange (pair energy for do_jtensor 1), anisotropy, DM, biquadratic,
    Zeeman per atom in mRy, as the columns Exc, Ani, DM, BQ, Zeeman of totenergy.*.out (their sum is the column Tot)."""
    H = ham_struct(S)
    N, M = S['Natom'], S['Mensemble']
    emomM = S['emomM'] if emomM is None else np.asfortranarray(emomM)
    t = np.zeros((5, M), order='F')
    lib().orc_energy_terms(C.byref(H), _p(emomM), _p(S['external_field']), _p(t))
    return t * (consts(S)['mub'] / consts(S)['mry'])


class SdState:
    """Mutable LLG state + work arrays for repeated orc_sd_step calls."""

    def __init__(self, S, sdealgh, delta_t, damping, temp=0.0, mompar=0, temprescale=1.0, red_atom_list=None, btorque=None):
        N, M = S['Natom'], S['Mensemble']
        # stt /= 'N': the spin-transfer-torque field btorque(3,N,M) the integrators add (midpoint.f90:86-97, depondt.f90:100-113)
        self.btorque = np.asfortranarray(btorque, dtype=np.float64) if btorque is not None else None
        self.S, self.sdealgh, self.delta_t, self.mompar, self.temprescale = S, sdealgh, delta_t, mompar, temprescale
        self.H = ham_struct(S)
        self.emom = S['emom'].copy(order='F')
        self.emomM = S['emomM'].copy(order='F')
        self.mmom = S['mmom'].copy(order='F')
        self.mmom0 = S['mmom0'].copy(order='F')
        self.lambda1 = np.full(N, damping) if np.isscalar(damping) else np.ascontiguousarray(damping, dtype=np.float64)
        self.temp = np.full(N, temp) if np.isscalar(temp) else np.ascontiguousarray(temp, dtype=np.float64)
        self.work = np.zeros(14 * N * M)
        # fixed-moment run: red_atom_list = the 1-based atoms that evolve (evolution.f90:38-44)
        self.frozen = None
        if red_atom_list is not None:
            self.frozen = np.ones(N, dtype=np.uint8)
            self.frozen[np.asarray(red_atom_list, dtype=np.int64) - 1] = 0

    def step(self, gauss=None):
        L = lib()
        S = self.S
        g = np.asfortranarray(gauss) if gauss is not None else None
        L.orc_set_btorque(_p(self.btorque))
        L.orc_sd_step(C.byref(self.H), self.sdealgh, _p(self.emom), _p(self.emomM), _p(self.mmom), _p(self.mmom0),
                      _p(S['external_field']), _p(S['Landeg']), _p(self.lambda1), _p(self.temp),
                      _d(self.temprescale), _d(self.delta_t), self.mompar, _p(g), _d(consts(self.S)['gama']),
                      _d(consts(self.S)['k_bolt']), _d(consts(self.S)['mub']), _d(consts(self.S)['mry']), _p(self.work),
                      _p(self.frozen))
        L.orc_set_btorque(None)

    def sum_moments(self):
        N, M = self.S['Natom'], self.S['Mensemble']
        m = np.zeros((3, M), order='F')
        lib().orc_sum_moments(N, M, _p(self.emomM), _p(m))
        return m


class Cumulants:
    """calc_and_print_cumulant (source/Measurement/prn_averages.f90:919-1034): weighted running means."""

    def __init__(self, natom):
        self.natom = natom
        self.cumuw = 0.0
        self.cumutotw = 0.0
        self.navrg = 0
        self.m1 = self.m2 = self.m4 = 0.0
        self.binder = 0.0

    def sample(self, msum):  # msum (3,M) = sum_i emomM
        for k in range(msum.shape[1]):
            m = msum[:, k]
            avrgme = np.sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]) / self.natom
            a2 = avrgme ** 2
            a4 = a2 ** 2
            self.cumuw += 1.0
            w, W = self.cumuw, self.cumutotw
            t1 = (self.m1 * W + avrgme * w) / (W + w)
            t2 = (self.m2 * W + a2 * w) / (W + w)
            t4 = (self.m4 * W + a4 * w) / (W + w)
            self.binder = 1 - (t4 / 3 / t2 ** 2)
            self.m1, self.m2, self.m4 = t1, t2, t4
            self.navrg += 1
            self.cumutotw += self.cumuw
        return self.m1, self.m2, self.m4, self.binder


def sd_run(S, inp, nstep=None, traj_atoms=(), want_rows=None, temp=0.0):
    """Replay of sd_mphase (source/sd_driver.f90:517-849): returns averages rows {iter: (mx,my,mz,m)},
    cumulant rows {sample_no: (m,m2,m4,U)}, trajectories {atom: {iter: (ex,ey,ez,m)}} and the final state.
    temp > 0: every step draws its 3*N*M normals from the reference's Ziggurat stream in its CURRENT state (rannum ->
    fill_rngarray, randomnumbers.f90:709-732), so a thermal initial phase run before continues into this phase as in
    the reference."""
    nstep = inp['nstep'] if nstep is None else nstep
    st = SdState(S, inp['sdealgh'], inp['timestep'], inp['damping'], temp=temp, mompar=inp['mompar'])
    N, M = S['Natom'], S['Mensemble']
    avg, cum, traj = {}, {}, {a: {} for a in traj_atoms}
    cu = Cumulants(N)
    avrg_step, cumu_step = inp['avrg_step'], inp['cumu_step']
    traj_step = inp.get('traj_step', 100)

    def measure(mstep):
        if inp['do_avrg'] == 'Y' and (mstep - 1) % avrg_step == 0:
            m = st.sum_moments()
            av = m / N
            nrm = np.sqrt((av ** 2).sum(axis=0))
            avg[mstep - 1] = (av[0].mean(), av[1].mean(), av[2].mean(), nrm.mean())
        for a in traj_atoms:
            if (mstep - 1) % traj_step == 0:
                traj[a][mstep - 1] = tuple(st.emom[:, a - 1, 0]) + (st.mmom[a - 1, 0],)
        if inp['do_cumu'] == 'Y' and mstep % cumu_step == 0:
            r = cu.sample(st.sum_moments())
            cum[cu.navrg // M] = r

    for mstep in range(1, nstep + 1):
        measure(mstep)
        if temp > 0.0:
            st.step(gauss=fill_rngarray(3 * N * M).reshape((3, N, M), order='F'))
        else:
            st.step()
    measure(nstep + 1)
    return dict(averages=avg, cumulants=cum, traj=traj, state=st)


# ---- reference RNG access (MT variant + ziggurat) ---------------------------------------------
def set_num_threads(n):
    """OpenMP threads of the restated loops (overrides OMP_NUM_THREADS); returns the count in force"""
    return int(lib().orc_set_num_threads(int(n)))


def rng_init(seed):
    lib().orc_rng_init(int(seed))


def rng_uniform(n):
    out = np.zeros(n)
    lib().orc_rng_uniform(_p(out), C.c_long(n))
    return out


def initmag1(S, seed):
    """Initmag 1 (magnetizationinit.f90:141-178): random directions from the reference's MT stream seeded with tseed
    (uppasd.f90:903-910); fills S['emom'] / S['emomM'] in place."""
    rng_init(seed)
    n1, n2, n3 = S['ncell']
    lib().orc_initmag1(S['Natom'], S['Mensemble'], S['NA'], n1, n2, n3, _p(S['mmom']), _p(S['emom']), _p(S['emomM']))
    return S


def zig_setup(seed):
    lib().orc_zig_setup(int(seed))


def fill_rngarray(n):
    out = np.zeros(n)
    lib().orc_fill_rngarray(_p(out), C.c_long(n))
    return out


# ---- thermal LLG and Monte Carlo replays (reference RNG streams) --------------------------------
def sd_run_thermal(S, sdealgh, delta_t, damping, temp, nstep, seed=1, sample_every=10, burn=0):
    """Thermal LLG with the reference's noise source: each step draws 3*N*M ziggurat normals in memory order
    (rannum -> fill_rngarray, randomnumbers.f90:709-732).  Returns per-sample |M|/N per ensemble."""
    zig_setup(seed)
    st = SdState(S, sdealgh, delta_t, damping, temp=temp)
    N, M = S['Natom'], S['Mensemble']
    out = []
    for step in range(1, nstep + 1):
        g = fill_rngarray(3 * N * M).reshape((3, N, M), order='F')
        st.step(gauss=g)
        if step > burn and step % sample_every == 0:
            m = st.sum_moments() / N
            out.append(np.sqrt((m ** 2).sum(axis=0)))
    return np.array(out), st


def mc_run(S, mode, temperature, nsweeps, seed=1, sample_every=1, burn=0, extfield=(0.0, 0.0, 0.0), init=True,
           reshuffle_every=None, before_sweep=None):
    """mc_mphase replay (source/mc_driver.f90:234-430): visiting order from choose_random_atom_x, redrawn every
    mcnstep/10 sweeps; per sweep the bulk draws of mc_evolve in the reference's order.  init=False continues the
    generators from their current state (a measurement phase that follows an initial phase)."""
    L = lib()
    if S.get('do_jtensor', 0) == 1:
        raise NotImplementedError('the Monte Carlo restatement covers scalar exchange only')
    if init:
        rng_init(seed)
        zig_setup(seed)
    N, M = S['Natom'], S['Mensemble']
    H = ham_struct(S)
    emom = S['emom'].copy(order='F')
    emomM = S['emomM'].copy(order='F')
    mmom = S['mmom'].copy(order='F')
    iflip = np.zeros(N, dtype=np.int32)
    L.orc_choose_random_atom_x(N, _p(iflip))
    ef = np.ascontiguousarray(extfield, dtype=np.float64)
    mags, ens = [], []
    for sweep in range(1, nsweeps + 1):
        if before_sweep is not None:
            before_sweep(sweep, emomM)          # where mc_mphase calls measure() and calc_energy: BEFORE sweep mcmstep
        fm = rng_uniform(3 * N * M)
        fg = fill_rngarray(3 * N * M)
        mf = rng_uniform(N * M) if mode == 'H' else None
        fa = rng_uniform(N * M)
        L.orc_mc_sweep(C.byref(H), C.c_char(mode.encode()), _p(iflip), _p(emomM), _p(emom), _p(mmom), _p(ef),
                       _p(S['external_field']), _d(temperature), _d(1.0), _p(fm), _p(fg), _p(mf), _p(fa),
                       _d(consts(S)['k_bolt']), _d(consts(S)['mub']))
        if sweep % (reshuffle_every or max(1, nsweeps // 10)) == 0:      # mcnstep/10 of the PHASE (mc_driver.f90:407-410)
            L.orc_choose_random_atom_x(N, _p(iflip))
        if sweep > burn and sweep % sample_every == 0:
            m = np.zeros((3, M), order='F')
            L.orc_sum_moments(N, M, _p(emomM), _p(m))
            mags.append(np.sqrt(((m / N) ** 2).sum(axis=0)))
            beff = np.zeros((3, N, M), order='F')
            e = L.orc_effective_field(C.byref(H), _p(emomM), _p(S['external_field']), _p(beff), None, None,
                                      _d(consts(S)['mub']), _d(consts(S)['mry']))
            ens.append(e / (N * M))
    return np.array(mags), np.array(ens), (emom, emomM, mmom)


class McState:
    """Mutable Monte Carlo state for replaying sweeps with EXTERNALLY supplied draws and visiting order (orc_mc_sweep is
    mc_evolve, montecarlo.f90:44-273, with the bulk draws as arguments).  Used by the deterministic chain-parity tests: the
    GPU's draws (asd_debug_mc_draws) and its sequential-equivalent visiting order (asd_get_mc_visit_order) go in, the chain
    must come out identical.  dm_quirk=False selects the consistent emomM form of the DM energy (see orc_set_dm_energy_quirk)."""

    def __init__(self, S, dm_quirk=True):
        self.S = S
        self.H = ham_struct(S)
        self.emom = S['emom'].copy(order='F')
        self.emomM = S['emomM'].copy(order='F')
        self.mmom = S['mmom'].copy(order='F')
        self.dm_quirk = dm_quirk

    def sweep(self, mode, temperature, order, u, g, extfield=(0.0, 0.0, 0.0), temprescale=1.0):
        """order: 1-based atoms in visiting order; u(4,N,M), g(3,N,M): per-ATOM draws (Metropolis: u[0:3] = trial-move
        uniforms, u[3] = acceptance; heat bath: u[0] = polar draw, u[1] = azimuth)."""
        L = lib()
        S = self.S
        N, M = S['Natom'], S['Mensemble']
        order = np.ascontiguousarray(order, dtype=np.int32)
        fm = np.asfortranarray(u[0:3])
        fg = np.asfortranarray(g)
        # mflip / flipprob_a are indexed by the POSITION in the sweep (montecarlo.f90:236,246), the GPU keys them by atom
        if mode == 'H':
            mf = np.asfortranarray(u[0][order - 1, :])
            fa = np.asfortranarray(u[1][order - 1, :])
        else:
            mf = None
            fa = np.asfortranarray(u[3][order - 1, :])
        ef = np.ascontiguousarray(extfield, dtype=np.float64)
        L.orc_set_dm_energy_quirk(1 if self.dm_quirk else 0)
        L.orc_mc_sweep(C.byref(self.H), C.c_char(mode.encode()), _p(order), _p(self.emomM), _p(self.emom), _p(self.mmom), _p(ef),
                       _p(S['external_field']), _d(temperature), _d(temprescale), _p(fm), _p(fg), _p(mf), _p(fa),
                       _d(consts(S)['k_bolt']), _d(consts(S)['mub']))
        L.orc_set_dm_energy_quirk(1)


# ---- topology (skyno T) ---------------------------------------------------------------------------
def delaunay_tri_tri(nx, ny, nz, nt):
    """delaunay_tri_tri (source/Measurement/topology.f90:307-380), loop for loop: simp(3, 2*nx*ny*nz*nt), 1-based."""
    def wrap_idx(x, y, z):
        return nx * ny * (z - 1) + nx * (y - 1) + x
    simp = []
    for z in range(1, nz + 1):
        for y in range(1, ny + 1):
            for x in range(1, nx + 1):
                for it in range(1, nt + 1):
                    simp.append((nt * wrap_idx(x, y, z) + it - nt,
                                 nt * wrap_idx(x % nx + 1, y, z) + it - nt,
                                 nt * wrap_idx(x, y % ny + 1, z) + it - nt))
    for z in range(1, nz + 1):
        for y in range(1, ny + 1):
            for x in range(1, nx + 1):
                for it in range(1, nt + 1):
                    simp.append((nt * wrap_idx(x, y, z) + it - nt,
                                 nt * wrap_idx(x, y % ny + 1, z) + it - nt,
                                 nt * wrap_idx((x - 2) % nx + 1, y % ny + 1, z) + it - nt))
    return np.asfortranarray(np.array(simp, dtype=np.int32).T)


def pontryagin_tri(emom, simp):
    """pontryagin_tri (topology.f90:78-116): sum over triangles and ensembles of 2 atan(m1.(m2 x m3) / (1 + m1.m2 + m1.m3 +
    m2.m3)) / (4 pi) / Mensemble, with f_volume(a, b, c) = (a x b) . c (math_functions.f90:59-73).  Also returns the
    per-ensemble sums / 4 pi."""
    emom = np.asarray(emom)
    M = emom.shape[2]
    per = np.zeros(M)
    for k in range(M):
        m1, m2, m3 = (emom[:, simp[c] - 1, k] for c in range(3))
        vol = (np.cross(m1.T, m2.T) * m3.T).sum(axis=1)
        d = 1.0 + (m1 * m2).sum(axis=0) + (m1 * m3).sum(axis=0) + (m2 * m3).sum(axis=0)
        per[k] = (2.0 * np.arctan(vol / d)).sum() / (4.0 * np.pi)
    return per.sum() / M, per


# ---- magnetic-field pulse (do_bpulse 1-4) -------------------------------------------------------------------------------
def bpulse_setup(do_bpulse, b0, step, par):
    """read_bpulse's derived constants ba, bb (fieldpulse.f90:196-209); par = bpulse_par(1:npar)"""
    bp = list(par) + [0.0] * (10 - len(par))
    ba = bb = 0.0
    if do_bpulse == 1:
        ba = (bp[0] - bp[1]) / np.log(bp[4] / bp[5])
        bb = (bp[2] - bp[3]) / np.log(bp[4] / bp[5])
    elif do_bpulse == 2:
        ba = 1.0
        bb = -1.0 / (2.0 * bp[2] ** 2)
    elif do_bpulse == 3:
        ba = bp[2] / (bp[1] - bp[0])
        bb = 1.0 / ((bp[1] - bp[0]) ** bp[2] * np.exp(-ba * bp[1]))
    return dict(do_bpulse=do_bpulse, b0=list(b0), step=int(step), bp=bp, ba=float(ba), bb=float(bb))


def bpulse_field(B, t):
    """bpulse (fieldpulse.f90:34-67) with exppulse / gaussianpulse / polexppulse / squarepulse (:73-119): bpulsefield(3) at time t"""
    bp, ba, bb, k = B['bp'], B['ba'], B['bb'], B['do_bpulse']
    if k == 1:
        if t <= bp[1]:
            tp = bp[5] * np.exp((t - bp[1]) / ba)
        elif t > bp[1] and t < bp[2]:
            tp = bp[5]
        else:
            tp = bp[5] * np.exp((bp[2] - t) / bb)
    elif k == 2:
        tp = bp[5] * ba * np.exp(bb * (t - bp[1]) ** 2)
    elif k == 3:
        tp = bp[5] * bb * (t - bp[0]) ** bp[2] * np.exp(-ba * t)
    elif k == 4:
        if t < bp[1]:
            tp = 0.0
        elif t >= bp[1] and t < bp[2]:
            tp = bp[5]
        else:
            tp = 0.0
    else:
        tp = 0.0
    return np.array([B['b0'][0] * tp, B['b0'][1] * tp, B['b0'][2] * tp])


def sd_run_bpulse(S, sdealgh, delta_t, damping, B, rstep, nstep):
    """sd_mphase with a field pulse (sd_driver.f90:389-393, 703-722, 770-779): the pulse is evaluated at delta_t * rstep before the
    first step and re-evaluated at delta_t * mstep after the moment update of every bpulse_step-th step; effective_field adds it
    to beff2 as time_external_field (hamiltonianactions.f90:241).  Returns the final SdState and the fields the steps saw."""
    st = SdState(S, sdealgh, delta_t, damping)
    ext0 = S['external_field'].copy(order='F')
    field = bpulse_field(B, delta_t * rstep)
    scount = 1
    seen = []
    for mstep in range(rstep + 1, rstep + nstep + 1):
        S['external_field'][...] = ext0 + field[:, None, None]
        seen.append(field.copy())
        st.step()
        if scount == B['step']:
            field = bpulse_field(B, delta_t * mstep)
            scount = 1
        else:
            scount += 1
    S['external_field'][...] = ext0
    return st, np.array(seen).T
