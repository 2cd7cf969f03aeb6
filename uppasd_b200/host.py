"""Host-side mirror of the reference drivers, in Python, for tests and benchmarks.

`Engine` wraps the explicit asd_* API with numpy arrays in the reference's Fortran shapes.
`FortranHost` plays the part of the reference's Fortran program around the LEGACY boundary: it owns the
module arrays (column-major numpy), calls fortrandata_set*_ / cudamdsim_*_ exactly like
`FortranData_Initiate` + `sd_mphaseCUDA` do (source/chelper.f90:166-186, source/sd_driver.f90:1118-1153) and
serves the measurement callbacks (`fortran_do_measurements`, `fortran_measure_moment`, ...).
"""
import ctypes as C

import numpy as np

from . import capi


class AsdError(RuntimeError):
    pass


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.asfortranarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def _i32(a):
    return np.asfortranarray(a, dtype=np.int32)


class Engine:
    """One engine = one GPU.  Arrays use the reference shapes: emom(3,N,M), mmom(N,M), nlist(z,N) 1-based."""

    def __init__(self, device=-1):
        self.lib = capi.load()
        h = C.c_void_p()
        self._chk(self.lib.asd_create(C.byref(h), device))
        self.h = h
        self.N = self.M = self.NH = 0
        self._keep = []

    def _chk(self, rc):
        if rc != 0:
            raise AsdError(self.lib.asd_last_error().decode())

    def close(self):
        if getattr(self, 'h', None):
            self.lib.asd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup -------------------------------------------------------------------------------------
    def set_constants(self, gama, k_bolt, mub, mry):
        self._chk(self.lib.asd_set_constants(self.h, gama, k_bolt, mub, mry))

    def set_system(self, natom, mensemble, nham, aham=None):
        self.N, self.M, self.NH = natom, mensemble, nham
        a = _i32(aham) if aham is not None else None
        self._chk(self.lib.asd_set_system(self.h, natom, mensemble, nham, _p(a)))

    def set_exchange(self, nlist, nlistsize, ncoup):
        nlist = _i32(nlist)
        self._chk(self.lib.asd_set_exchange(self.h, nlist.shape[0], _p(nlist), _p(_i32(nlistsize)), _p(_f64(ncoup))))

    def set_jtensor(self, nlist, nlistsize, j_tens):
        """j_tens: (9, z, nHam) = the Fortran (3,3,z,nHam) array"""
        nlist = _i32(nlist)
        self._chk(self.lib.asd_set_jtensor(self.h, nlist.shape[0], _p(nlist), _p(_i32(nlistsize)), _p(_f64(j_tens))))

    def set_dm(self, dmlist, dmlistsize, dm_vect):
        dmlist = _i32(dmlist)
        self._chk(self.lib.asd_set_dm(self.h, dmlist.shape[0], _p(dmlist), _p(_i32(dmlistsize)), _p(_f64(dm_vect))))

    def set_bq(self, bqlist, bqlistsize, j_bq):
        bqlist = _i32(bqlist)
        self._chk(self.lib.asd_set_bq(self.h, bqlist.shape[0], _p(bqlist), _p(_i32(bqlistsize)), _p(_f64(j_bq))))

    def set_lattice_hint(self, na, ncell, bc):
        self._chk(self.lib.asd_set_lattice_hint(self.h, na, ncell[0], ncell[1], ncell[2], ''.join(bc).encode()))

    def set_anisotropy(self, taniso, eaniso, kaniso, sb):
        self._chk(self.lib.asd_set_anisotropy(self.h, _p(_i32(taniso)), _p(_f64(eaniso)), _p(_f64(kaniso)), _p(_f64(sb))))

    def set_external_field(self, ext):
        self._chk(self.lib.asd_set_external_field(self.h, _p(_f64(ext, (3, self.N, self.M)))))

    def set_torque(self, bt):
        self._chk(self.lib.asd_set_torque(self.h, _p(_f64(bt, (3, self.N, self.M)))))

    def set_time_field(self, first_step, tfield):
        """tfield(3, nsteps): uniform time-dependent field of the steps first_step ... (None clears)"""
        if tfield is None:
            self._chk(self.lib.asd_set_time_field(self.h, 0, 0, None))
            return
        f = _f64(tfield)
        assert f.ndim == 2 and f.shape[0] == 3, f.shape
        self._chk(self.lib.asd_set_time_field(self.h, first_step, f.shape[1], _p(f)))

    def set_llg(self, sdealgh, delta_t, landeg=None, lambda1=None, temp=None, temprescale=1.0, mompar=0, seed=20261017):
        def arr(x):
            if x is None:
                return None
            return np.full(self.N, float(x)) if np.isscalar(x) else np.ascontiguousarray(x, dtype=np.float64)
        a, b, c = arr(landeg), arr(lambda1), arr(temp)
        self._chk(self.lib.asd_set_llg(self.h, sdealgh, delta_t, _p(a), _p(b), _p(c), temprescale, mompar, seed))

    def set_evolving_atoms(self, red_atom_list=None):
        """red_atom_list: 1-based atoms that evolve (fixed-moment runs); None: all"""
        if red_atom_list is None:
            self._chk(self.lib.asd_set_evolving_atoms(self.h, 0, None))
        else:
            r = _i32(red_atom_list)
            self._chk(self.lib.asd_set_evolving_atoms(self.h, int(r.size), _p(r)))

    def set_moments(self, emom, mmom, mmom0=None):
        e, m = _f64(emom, (3, self.N, self.M)), _f64(mmom, (self.N, self.M))
        m0 = _f64(mmom0, (self.N, self.M)) if mmom0 is not None else None
        self._chk(self.lib.asd_set_moments(self.h, _p(e), _p(m), _p(m0)))

    def commit(self):
        self._chk(self.lib.asd_commit(self.h))

    # ---- compute -----------------------------------------------------------------------------------
    def get_moments(self, out=None):
        """out: optional (emom, emomM, mmom) Fortran-ordered arrays to fill (e.g. views of pinned host memory); an entry that is
        None is not copied (the C ABI takes NULL for it)"""
        if out is None:
            emom = np.zeros((3, self.N, self.M), order='F')
            emomM = np.zeros((3, self.N, self.M), order='F')
            mmom = np.zeros((self.N, self.M), order='F')
        else:
            emom, emomM, mmom = out
        self._chk(self.lib.asd_get_moments(self.h, _p(emom), _p(emomM), _p(mmom)))
        return emom, emomM, mmom

    def effective_field(self, parts=False, energy=True):
        beff = np.zeros((3, self.N, self.M), order='F')
        b1 = np.zeros((3, self.N, self.M), order='F') if parts else None
        b2 = np.zeros((3, self.N, self.M), order='F') if parts else None
        en = np.zeros(self.M) if energy else None
        self._chk(self.lib.asd_effective_field(self.h, _p(beff), _p(b1), _p(b2), _p(en)))
        return (beff, b1, b2, en) if parts else (beff, en)

    def sd_steps(self, nsteps, first_step=1):
        self._chk(self.lib.asd_sd_steps(self.h, nsteps, first_step))

    def sd_run(self, nsteps, first_step=1, sample_every=1, out=None):
        """the measurement-phase loop in one call (asd_sd_run): msum(3, M, nsamples), sampled after every sample_every-th step;
        one device-to-host copy and one synchronisation at the end.  out: optional Fortran-ordered landing array"""
        ns = nsteps // sample_every if sample_every > 0 else 1
        msum = out if out is not None else np.zeros((3, self.M, max(ns, 1)), order='F')
        got = C.c_long(0)
        self._chk(self.lib.asd_sd_run(self.h, nsteps, first_step, sample_every, _p(msum), C.byref(got)))
        return msum[:, :, :got.value]

    def mc_sweeps(self, mode, nsweeps, temperature, first_sweep=1, temprescale=1.0, extfield=None):
        ef = np.ascontiguousarray(extfield, dtype=np.float64) if extfield is not None else None
        self._chk(self.lib.asd_mc_sweeps(self.h, mode.encode(), nsweeps, first_sweep, temperature, temprescale, _p(ef)))

    def set_mc_layout(self, layout):
        self._chk(self.lib.asd_set_mc_layout(self.h, layout))

    def mc_colouring(self):
        """(layout, ncolours, period): what the next MC sweep uses"""
        lay, nc = C.c_int(), C.c_int()
        per = (C.c_int * 3)()
        self._chk(self.lib.asd_mc_colouring(self.h, C.byref(lay), C.byref(nc), per))
        return lay.value, nc.value, tuple(per)

    def get_mc_colours(self):
        col = np.full(self.N, -1, dtype=np.int32)
        self._chk(self.lib.asd_get_mc_colours(self.h, _p(col)))
        return col

    def get_mc_visit_order(self):
        """1-based atoms in a sequential visiting order that reproduces the chain of the next sweep"""
        order = np.zeros(self.N, dtype=np.int32)
        self._chk(self.lib.asd_get_mc_visit_order(self.h, _p(order)))
        return order

    def debug_mc_draws(self, sweep):
        """(u(4,N,M), g(3,N,M)): the draws of one sweep as the update kernels compute them (test hook)"""
        u = np.zeros((4, self.N, self.M), order='F')
        g = np.zeros((3, self.N, self.M), order='F')
        self._chk(self.lib.asd_debug_mc_draws(self.h, sweep, _p(u), _p(g)))
        return u, g

    def measure(self, energy=False):
        msum = np.zeros((3, self.M), order='F')
        en = np.zeros(self.M) if energy else None
        self._chk(self.lib.asd_measure(self.h, _p(msum), _p(en)))
        return (msum, en) if energy else msum

    def energy_terms(self):
        """(5, M): exchange, anisotropy, DM, biquadratic, Zeeman energy per atom in mRy (energy.f90 estimators)"""
        out = np.zeros((5, self.M), order='F')
        self._chk(self.lib.asd_energy_terms(self.h, _p(out)))
        return out

    def measure_sublattice(self, na):
        """(3, NA, M): sums of emomM per basis atom (buffer_proj_avrg)"""
        out = np.zeros((3, na, self.M), order='F')
        self._chk(self.lib.asd_measure_sublattice(self.h, na, _p(out)))
        return out

    def set_triangulation(self, simp):
        """simp(3, nsimp): 1-based corner atoms of the triangles (delaunay_tri_tri)"""
        simp = np.asfortranarray(simp, dtype=np.int32)
        self._chk(self.lib.asd_set_triangulation(self.h, simp.shape[1], _p(simp)))

    def skyrmion_number(self):
        """per ensemble: sum of the triangles' solid angles / 4 pi (pontryagin_tri before the ensemble mean)"""
        q = np.zeros(self.M)
        self._chk(self.lib.asd_skyrmion_number(self.h, _p(q)))
        return q

    def get_atoms(self, atoms):
        """(4, n, M): ex, ey, ez, |m| of the 1-based atoms"""
        a = np.ascontiguousarray(atoms, dtype=np.int32)
        out = np.zeros((4, len(a), self.M), order='F')
        self._chk(self.lib.asd_get_atoms(self.h, len(a), _p(a), _p(out)))
        return out

    def time_sd_steps(self, nsteps, first_step=1, stages=False):
        tot = C.c_float(0)
        st = (C.c_float * 2)(0, 0)
        self._chk(self.lib.asd_time_sd_steps(self.h, nsteps, first_step, C.byref(tot), st if stages else None))
        return (tot.value, (st[0], st[1])) if stages else tot.value

    def time_mc_sweeps(self, mode, nsweeps, temperature):
        tot = C.c_float(0)
        self._chk(self.lib.asd_time_mc_sweeps(self.h, mode.encode(), nsweeps, temperature, C.byref(tot)))
        return tot.value

    def layout_info(self):
        """dict(staged, runs, ucap, union, tile_slots, extra_staged, planes): the field path of the LLG stage kernels"""
        info = (C.c_int * 6)()
        self._chk(self.lib.asd_layout_info(self.h, info))
        return dict(staged=info[0], runs=info[1], ucap=info[2], union=info[3], tile_slots=info[4], extra_staged=info[5] & 1,
                    planes=(info[5] >> 1) & 1)

    def launch_count(self):
        return self.lib.asd_launch_count(self.h)

    def synchronize(self):
        self._chk(self.lib.asd_synchronize(self.h))

    # ---- on-device tables --------------------------------------------------------------------------
    def build_lattice_table(self, kind, na, ncell, bc, nslot, cell_atom, cell_shift, coupling):
        nslot = np.ascontiguousarray(nslot, dtype=np.int32)
        cell_atom = np.ascontiguousarray(cell_atom, dtype=np.int32)      # (NA, maxslot)
        cell_shift = np.ascontiguousarray(cell_shift, dtype=np.int32)    # (NA, maxslot, 3)
        coupling = np.ascontiguousarray(coupling, dtype=np.float64)      # (NA, maxslot, ncomp)
        maxslot = cell_atom.shape[1]
        self._chk(self.lib.asd_build_lattice_table(self.h, kind, na, ncell[0], ncell[1], ncell[2],
                                                   ''.join(bc).encode(), maxslot, _p(nslot), _p(cell_atom),
                                                   _p(cell_shift), _p(coupling)))

    def get_table(self, kind):
        z, nc = C.c_int(0), C.c_int(0)
        self._chk(self.lib.asd_get_table_dims(self.h, kind, C.byref(z), C.byref(nc)))
        lst = np.zeros((z.value, self.N), dtype=np.int32, order='F')
        size = np.zeros(self.NH, dtype=np.int32)
        coup = np.zeros((nc.value, z.value, self.NH), order='F')
        self._chk(self.lib.asd_get_table(self.h, kind, _p(lst), _p(size), _p(coup)))
        return lst, size, (coup[0] if nc.value == 1 else coup)

    def init_moments_tilted(self, amplitude, mmom_basis):
        mb = np.ascontiguousarray(mmom_basis, dtype=np.float64)
        self._chk(self.lib.asd_init_moments_tilted(self.h, amplitude, len(mb), _p(mb)))


    # ---- multi-GPU -------------------------------------------------------------------------------------
    def set_ensemble_offset(self, first_ensemble):
        self._chk(self.lib.asd_set_ensemble_offset(self.h, first_ensemble))

    def set_slab(self, nslabs, slab_index, halo_planes):
        self._chk(self.lib.asd_set_slab(self.h, nslabs, slab_index, halo_planes))

    def slab_export(self):
        buf = C.create_string_buffer(self.lib.asd_slab_handle_bytes())
        self._chk(self.lib.asd_slab_export(self.h, buf))
        return buf.raw

    def slab_connect_ipc(self, lower, upper):
        self._chk(self.lib.asd_slab_connect_ipc(self.h, C.c_char_p(lower), C.c_char_p(upper)))

    def slab_connect_local(self, lower, upper):
        self._chk(self.lib.asd_slab_connect_local(self.h, lower.h, upper.h))

    def slab_status(self):
        ep, err = C.c_ulonglong(0), C.c_int(0)
        self._chk(self.lib.asd_slab_status(self.h, C.byref(ep), C.byref(err)))
        return ep.value, err.value


def engine_from_system(S, consts, sdealgh=1, delta_t=1e-16, damping=0.05, temp=0.0, mompar=0, seed=20261017,
                       device=-1, lattice_hint=None):
    """Feeds a system dict holding reference-shaped tables (nlist, ncoup, ... as the Fortran host has them) to a new Engine.
    lattice_hint = (NA, ncell, bc): announce the supercell shape (asd_set_lattice_hint)."""
    e = Engine(device)
    e.set_constants(consts['gama'], consts['k_bolt'], consts['mub'], consts['mry'])
    e.set_system(S['Natom'], S['Mensemble'], S['nHam'], S['aHam'])
    if lattice_hint is not None:
        e.set_lattice_hint(*lattice_hint)
    ex = S['exchange']
    if S.get('do_jtensor', 0) == 1:
        e.set_jtensor(ex['list'], ex['listsize'], ex['coup'])
    else:
        e.set_exchange(ex['list'], ex['listsize'], ex['coup'])
    if S.get('dm') is not None:
        e.set_dm(S['dm']['list'], S['dm']['listsize'], S['dm']['coup'])
    if S.get('bq') is not None:
        e.set_bq(S['bq']['list'], S['bq']['listsize'], S['bq']['coup'])
    if S.get('aniso') is not None:
        a = S['aniso']
        e.set_anisotropy(a['taniso'], a['eaniso'], a['kaniso'], a['sb'])
    e.set_external_field(S['external_field'])
    e.set_llg(sdealgh, delta_t, landeg=S['Landeg'], lambda1=damping, temp=temp, mompar=mompar, seed=seed)
    e.set_moments(S['emom'], S['mmom'], S['mmom0'])
    e.commit()
    return e


class FortranHost:
    """Plays the reference's Fortran program around the legacy boundary (sd_mphaseCUDA)."""

    def __init__(self, S, consts, sdealgh, nstep, delta_t, damping, temp=0.0, mompar=0, rstep=0, gpu_rng_seed=1,
                 avrg_step=100, cumu_step=50, do_avrg='Y', do_cumu='N', lattice=None):
        """lattice = (NA, (N1, N2, N3), 'PPP'): also announce the supercell shape (fortrandata_setlattice_)"""
        self.lib = capi.load()
        self.S = S
        self.lattice = None
        if lattice is not None:
            na, nc, bc = lattice
            self.lattice = [C.c_uint(na), C.c_uint(nc[0]), C.c_uint(nc[1]), C.c_uint(nc[2])] + [C.c_char(b.encode()) for b in bc]
        N, M = S['Natom'], S['Mensemble']
        self.N, self.M = N, M
        ci = lambda v: C.c_int(v)
        cu = lambda v: C.c_uint(v)
        cd = lambda v: C.c_double(v)
        # scalars live in ctypes objects (Fortran module variables)
        self.sc = dict(stt=C.c_char(b'N'), SDEalgh=ci(sdealgh), rstep=cu(rstep), nstep=cu(nstep), Natom=cu(N),
                       Mensemble=cu(M), max_no_neigh=cu(S['exchange']['z']), delta_t=cd(delta_t),
                       gamma=cd(consts['gama']), k_bolt=cd(consts['k_bolt']), mub=cd(consts['mub']),
                       damping=cd(damping), binderc=cd(0.0), mavg=cd(0.0), mompar=ci(mompar),
                       initexc=C.c_char(b'N'), do_dm=cu(1 if S.get('dm') is not None else 0),
                       max_no_dmneigh=cu(S['dm']['z'] if S.get('dm') is not None else 1), do_jtensor=cu(1 if S.get('do_jtensor', 0) == 1 else 0),
                       do_aniso=cu(1 if S.get('aniso') is not None else 0), nHam=cu(S['nHam']),
                       gpu_mode=ci(1), gpu_rng=ci(0), gpu_rng_seed=ci(gpu_rng_seed))
        z = np.zeros
        jt = S.get('do_jtensor', 0) == 1
        self.arr = dict(
            ncoup=_f64(S['exchange']['coup']) if not jt else z((1, 1), order='F'), nlist=_i32(S['exchange']['list']),
            nlistsize=_i32(S['exchange']['listsize']),
            beff=z((3, N, M), order='F'), b2eff=z((3, N, M), order='F'), emomM=S['emomM'].copy(order='F'),
            emom=S['emom'].copy(order='F'), emom2=z((3, N, M), order='F'), external_field=_f64(S['external_field']),
            mmom=S['mmom'].copy(order='F'), btorque=z((3, N, M), order='F'), Temp_array=np.full(N, float(temp)),
            mmom0=S['mmom0'].copy(order='F'), mmom2=z((N, M), order='F'), mmomi=S['mmomi'].copy(order='F'),
            dm_vect=_f64(S['dm']['coup']) if S.get('dm') is not None else z((3, 1, 1), order='F'),
            dmlist=_i32(S['dm']['list']) if S.get('dm') is not None else z((1, 1), dtype=np.int32),
            dmlistsize=_i32(S['dm']['listsize']) if S.get('dm') is not None else z(1, dtype=np.int32),
            j_tens=_f64(S['exchange']['coup']) if jt else z((3, 3, 1, 1), order='F'),
            kaniso=_f64(S['aniso']['kaniso']) if S.get('aniso') is not None else z((2, 1), order='F'),
            eaniso=_f64(S['aniso']['eaniso']) if S.get('aniso') is not None else z((3, 1), order='F'),
            taniso=_i32(S['aniso']['taniso']) if S.get('aniso') is not None else z(1, dtype=np.int32),
            sb=_f64(S['aniso']['sb']) if S.get('aniso') is not None else z(1), aHam=_i32(S['aHam']),
            Landeg=_f64(S['Landeg']))
        self.avrg_step, self.cumu_step, self.do_avrg, self.do_cumu = avrg_step, cumu_step, do_avrg, do_cumu
        self.averages = {}      # iter -> (mx,my,mz,m)
        self.samples = []       # (mstep, emom copy) for every measured step
        self.flushed_at = None
        self._cbs = (capi.CB_DO(self._do_measurements), capi.CB_MEASURE(self._measure_moment),
                     capi.CB_FLUSH(self._flush), capi.CB_STATUS(self._status))

    # --- callbacks (chelper.f90:73-160) ---
    def _do_measurements(self, mstep, do_copy):
        ms = mstep[0]
        copy = 0
        if self.do_avrg == 'Y' and (ms - 1) % self.avrg_step == 0:
            copy = 1
        if self.do_cumu == 'Y' and ms % self.cumu_step == 0:
            copy = 1
        do_copy[0] = copy

    def _measure_moment(self, emomM, emom, mmom, mstep):
        ms = mstep[0]
        N, M = self.N, self.M
        eM = np.ctypeslib.as_array(emomM, shape=(M, N, 3))
        if self.do_avrg == 'Y' and (ms - 1) % self.avrg_step == 0:
            av = eM.sum(axis=1) / N          # (M,3)
            nrm = np.sqrt((av ** 2).sum(axis=1))
            self.averages[ms - 1] = (av[:, 0].mean(), av[:, 1].mean(), av[:, 2].mean(), nrm.mean())
        self.samples.append(ms)

    def _flush(self, mstep):
        self.flushed_at = mstep[0]

    def _status(self, mavg):
        eM = self.arr['emomM']
        m = eM.sum(axis=1) / self.N
        mavg[0] = float(np.sqrt((m ** 2).sum(axis=0)).mean())

    # --- the call sequence of FortranData_Initiate + sd_mphaseCUDA ---
    def initiate(self):
        s, a, L = self.sc, self.arr, self.lib
        r = lambda k: C.cast(C.byref(s[k]), C.c_void_p)
        L.asd_set_callbacks(*self._cbs)
        L.fortrandata_setconstants_(r('stt'), r('SDEalgh'), r('rstep'), r('nstep'), r('Natom'), r('Mensemble'),
                                    r('max_no_neigh'), r('delta_t'), r('gamma'), r('k_bolt'), r('mub'), r('damping'),
                                    r('binderc'), r('mavg'), r('mompar'), r('initexc'), r('do_dm'),
                                    r('max_no_dmneigh'), r('do_jtensor'), r('do_aniso'), r('nHam'))
        L.fortrandata_setmatrices_(*[_p(a[k]) for k in ('ncoup', 'nlist', 'nlistsize', 'beff', 'b2eff', 'emomM', 'emom',
                                                        'emom2', 'external_field', 'mmom', 'btorque', 'Temp_array',
                                                        'mmom0', 'mmom2', 'mmomi', 'dm_vect', 'dmlist', 'dmlistsize',
                                                        'j_tens', 'kaniso', 'eaniso', 'taniso', 'sb', 'aHam')])
        L.fortrandata_setinputdata_(r('gpu_mode'), r('gpu_rng'), r('gpu_rng_seed'))
        bq = self.S.get('bq')
        if bq is not None:
            # the extras a maintainer passes for the biquadratic table (ham%bqlist, ham%bqlistsize, ham%j_bq, nn_bq_tot)
            self._bq = dict(do_bq=C.c_uint(1), nn=C.c_uint(bq['z']), lst=_i32(bq['list']), size=_i32(bq['listsize']), j=_f64(bq['coup']))
            q = self._bq
            L.fortrandata_setextras_(_p(a['Landeg']), None, None, C.cast(C.byref(q['do_bq']), C.c_void_p),
                                     C.cast(C.byref(q['nn']), C.c_void_p), _p(q['lst']), _p(q['size']), _p(q['j']))
        else:
            L.fortrandata_setextras_(_p(a['Landeg']), None, None, None, None, None, None, None)
        if self.lattice is not None:
            L.fortrandata_setlattice_(*[C.cast(C.byref(x), C.c_void_p) for x in self.lattice])
        else:
            L.fortrandata_setlattice_(None, None, None, None, None, None, None)
        L.cudamdsim_initiateconstants_()
        L.cudamdsim_initiatematrices_()
        return self

    def run(self):
        self.initiate()
        self.lib.cudamdsim_measurementphase_()
        return self

    def layout_info(self):
        """field path of the engine behind the legacy boundary (Engine.layout_info)"""
        info = (C.c_int * 6)()
        h = self.lib.asd_legacy_engine()
        if not h or self.lib.asd_layout_info(h, info):
            raise AsdError(self.lib.asd_last_error().decode())
        return dict(staged=info[0], runs=info[1], ucap=info[2], union=info[3], tile_slots=info[4], extra_staged=info[5])

    # --- the new sibling entries, called the way sd_iphase / mc_mphase would ---
    def initial_phase(self, nstep, temp, delta_t, damping, sdealgh, first_step=1):
        b = C.byref
        self.lib.cudamdsim_initialphase_(b(C.c_uint(nstep)), b(C.c_double(temp)), b(C.c_double(delta_t)), b(C.c_double(damping)),
                                         b(C.c_int(sdealgh)), b(C.c_uint(first_step)))

    def mc_evolve(self, mode, nsweeps, temp, first_sweep=1, extfield=(0.0, 0.0, 0.0), upload=False, temprescale=1.0):
        b = C.byref
        ef = (C.c_double * 3)(*extfield)
        self.lib.cudamcsim_evolve_(b(C.c_char(mode.encode())), b(C.c_uint(nsweeps)), b(C.c_uint(first_sweep)), b(C.c_double(temp)),
                                   b(C.c_double(temprescale)), ef, b(C.c_int(1 if upload else 0)))

    # --- pyasd's entry points (source/pyasd.f90), called the way the Python package uppasd calls them ---
    def relax(self, mode, nstep, temperature, timestep, damping):
        """relax_: returns moments(3, N, M) = emomM"""
        b = C.byref
        mom = np.zeros((3, self.N, self.M), order='F')
        self.lib.relax_(_p(mom), b(C.c_int(self.N)), b(C.c_int(self.M)), b(C.c_char(mode.encode())), b(C.c_int(nstep)),
                        b(C.c_double(temperature)), b(C.c_double(timestep)), b(C.c_double(damping)))
        return mom

    def get_emom(self):
        mom = np.zeros((3, self.N, self.M), order='F')
        self.lib.get_emom_(_p(mom), C.byref(C.c_int(self.N)), C.byref(C.c_int(self.M)))
        return mom

    def put_emom(self, moments):
        m = _f64(moments, (3, self.N, self.M))
        self.lib.put_emom_(_p(m), C.byref(C.c_int(self.N)), C.byref(C.c_int(self.M)))

    def get_beff(self):
        f = np.zeros((3, self.N, self.M), order='F')
        self.lib.get_beff_(_p(f), C.byref(C.c_int(self.N)), C.byref(C.c_int(self.M)))
        return f

    def get_energy(self):
        en = C.c_double(0.0)
        self.lib.get_energy_(C.byref(en))
        return en.value
