"""Driver-level mirror of the reference around the hot path: what `sd` does for a run directory, for the slice of
inpsd.dat this path serves, with every per-step loop body replaced by calls into libuppasd_b200.so.

  setup            uppasd.f90:556 setup_simulation (geometry.f90:337-488, magnetizationinit.f90:234-273,
                   hamiltonianinit.f90:857-976; tables through the on-device builder)
  initial phase    sd_driver.f90:42-291 sd_iphase, mc_driver.f90:45-223 mc_iphase
  measurement      sd_driver.f90:300-850 sd_mphase, mc_driver.f90:234-430 mc_mphase: measure() BEFORE the step, rows
                   labelled mstep-1, averages when mod(mstep-1,avrg_step)==0, cumulants when mod(mstep,cumu_step)==0,
                   energy on the averages cadence when plotenergy>0, restart file with every averages flush
  relax            pyasd.f90:255-298 relax_ -> sd_minimal (sd_driver.f90:1162) / mc_minimal (mc_driver.f90:457)

Only what the hot path needs is here; keywords outside it are ignored, features outside it are refused loudly.
"""
import os
import warnings

import numpy as np

from . import alloy, asdio, fields, host, lattice, observables, refrng

# source/Parameters/constants.f90:14-29
CONSTANTS = dict(gama=1.760859644e11, k_bolt=1.38064852e-23, mub=9.274009994e-24, mry=2.179872325e-21)


class Unsupported(RuntimeError):
    pass


class Simulation:
    def __init__(self, inp, directory='.', device=-1, seed=None, consts=None):
        """inp: dict from asdio.read_inpsd (or a path to inpsd.dat)"""
        if isinstance(inp, str):
            directory = os.path.dirname(os.path.abspath(inp))
            inp = asdio.read_inpsd(inp)
        self.inp, self.dir = inp, directory
        self.c = dict(consts or CONSTANTS)
        if inp['aunits'] == 'Y':                       # inputhandler.f90:1685-1704
            self.c = dict(gama=1.0, k_bolt=1.0, mub=1.0, mry=1.0)
        if inp.get('unserved'):
            raise Unsupported('inpsd.dat switches on features outside the hot path served here: '
                              + ', '.join('%s %s' % kv for kv in inp['unserved']))
        for key in inp.get('unwritten', []):
            warnings.warn('inpsd.dat asks for %r: that measurement is not written by this driver (the dynamics are unaffected)' % key)
        for key in inp.get('ignored', []):
            warnings.warn('inpsd.dat keyword %r is not known to this driver and was ignored' % key)
        if inp['map_multiple']:
            # the device table builder drops a second coupling between the same pair like the reference does WITHOUT map_multiple
            # (hamiltonianinit.f90:1059); keeping duplicates is not implemented, so refuse rather than mount shorter lists
            raise Unsupported('map_multiple T (several couplings between one pair of atoms) is not served')
        if inp['do_ralloy'] not in (0, 1):
            raise Unsupported('do_ralloy %d' % inp['do_ralloy'])
        if inp['do_ralloy'] == 1:
            # random alloy: occupancy from the reference's generator, one coupling row per atom (uppasd_b200/alloy.py)
            if inp.get('dm') or inp.get('bq') or inp.get('anisotropy') or inp['do_jtensor'] == 1:
                raise Unsupported('do_ralloy 1 is served for scalar exchange only (no dm / bq / anisotropy / do_jtensor)')
            if inp['do_reduced'] == 'Y':
                raise Unsupported('do_ralloy 1 needs one Hamiltonian row per atom (do_reduced N)')
            if inp['initmag'] != 3:
                raise Unsupported('do_ralloy 1: initmag %d (only 3, moments from the momfile)' % inp['initmag'])
        if inp['do_jtensor'] == 1 and (inp['mode'] != 'S' or inp['ip_mode'] not in ('N', 'S')):
            raise Unsupported('do_jtensor 1 is served for spin dynamics only (mode / ip_mode S)')
        if inp['mode'] not in ('S', 'M', 'H') or inp['ip_mode'] not in ('N', 'S', 'M', 'H'):
            raise Unsupported('mode %s / ip_mode %s: only S (spin dynamics), M (Metropolis), H (heat bath)' % (inp['mode'], inp['ip_mode']))
        self.seed = int(seed if seed is not None else (inp['gpu_rng_seed'] or inp['tseed']))
        self.out = asdio.OutputFiles(directory, inp['simid'])
        self.rstep = 0
        self._setup(device)

    # ------------------------------------------------------------------------------------------------
    def _setup_alloy(self, device):
        """do_ralloy 1 (fully occupied supercell): neighbour lists of the underlying lattice built on the device, the shell of
        every list entry carried through the builder in place of a coupling, then the chemistry-dependent couplings mounted
        per atom (alloy.mount) and handed to a second engine together with the supercell shape"""
        inp, c = self.inp, self.c
        cell = np.asarray(inp['cell'], dtype=float)
        bas, atype_inp, nch, chconc = asdio.read_posfile_alloy(inp['posfile'], cell, inp['posfiletype'])
        bas = lattice.fold_basis(cell, bas)
        na, nchmax = chconc.shape
        if np.abs(np.array([chconc[i, :nch[i]].sum() for i in range(na)]) - 1.0).max() > 1e-9:
            raise Unsupported('do_ralloy 1: dilute alloys (site concentrations that do not add up to 1) are not served')
        n1, n2, n3 = inp['ncell']
        natom, mens = na * n1 * n2 * n3, inp['mensemble']
        self.na, self.natom, self.mens = na, natom, mens
        ammom, aemom, landeg = asdio.read_momfile_alloy(inp['momfile'], na, nchmax, inp['landeg_glob'])
        self.anumb = (np.arange(natom, dtype=np.int32) % na) + 1
        self.atype = atype_inp[self.anumb - 1]
        self.achtype = alloy.occupancy(na, (n1, n2, n3), nch, chconc, inp['tseed'])
        if (self.achtype == 0).any():
            raise Unsupported('do_ralloy 1: the concentrations leave sites vacant for this supercell size')
        idx = np.arange(natom)
        i0, ix, iy, iz = idx % na, (idx // na) % n1, (idx // (na * n1)) % n2, idx // (na * n1 * n2)
        self.coord = (np.outer(cell[0], ix) + np.outer(cell[1], iy) + np.outer(cell[2], iz)) + bas[:, i0]
        if inp['do_prnstruct'] in (1, 2, 4):
            self.out.coord(self.coord, self.atype, self.anumb)
        nn, red, xc, nntype = asdio.read_pairfile_alloy(inp['exchange'], atype_inp, nchmax, bas, cell, inp['maptype'], inp['posfiletype'])
        ns, ca, cs, sh = lattice.stencil(cell, bas, atype_inp, nn, red, inp['sym'], nntype, ncell=(n1, n2, n3))
        # pass 1: the lattice's neighbour lists, every entry tagged with its shell (the "coupling" of the builder)
        t = host.Engine(device)
        t.set_constants(c['gama'], c['k_bolt'], c['mub'], c['mry'])
        t.set_system(natom, 1, natom, None)
        t.build_lattice_table(0, na, (n1, n2, n3), inp['bc'], ns, ca, cs, np.asarray(sh, dtype=np.float64)[:, :, None] + 1.0)
        nlist, nlistsize, tag = t.get_table(0)
        t.close()
        shell = np.clip(np.rint(tag).astype(np.int64) - 1, 0, None)
        ncoup = alloy.mount(nlist, nlistsize, shell, self.atype, self.anumb, self.achtype, xc, ammom, c['mry'], c['mub'])
        self.tables = dict(nlist=nlist, nlistsize=nlistsize, ncoup=ncoup)
        # pass 2: the engine of the run, one coupling row per atom, atoms in brick order through the supercell hint
        e = host.Engine(device)
        e.set_constants(c['gama'], c['k_bolt'], c['mub'], c['mry'])
        e.set_system(natom, mens, natom, None)
        e.set_lattice_hint(na, (n1, n2, n3), inp['bc'])
        e.set_exchange(nlist, nlistsize, ncoup)
        site, chem = self.anumb - 1, self.achtype - 1
        self.landeg = 0.5 * landeg[site, chem]
        self.engine = e
        self._set_field(inp['ip_hfield'] if inp['ip_mode'] != 'N' else inp['hfield'])
        self._llg(inp['sdealgh'], inp['timestep'], inp['damping'], inp['temp'])
        e.commit()
        mmom = np.asfortranarray(np.repeat(np.abs(ammom[site, chem])[:, None], mens, axis=1))
        if inp['initmag'] == 3:
            emom = np.asfortranarray(np.repeat(aemom[:, site, chem][:, :, None], mens, axis=2))
        else:
            # Initmag 1 continues the generator that dealt the species (same stream, magnetizationinit.f90:141-178): not restated
            raise Unsupported('do_ralloy 1 with initmag 1')
        self.mmom0 = mmom.copy(order='F')
        e.set_moments(emom, mmom, self.mmom0)
        self.atype_cell = atype_inp

    def _setup(self, device):
        if self.inp['do_ralloy'] == 1:
            return self._setup_alloy(device)
        inp, c = self.inp, self.c
        cell = np.asarray(inp['cell'], dtype=float)
        bas, atype_inp = asdio.read_posfile(inp['posfile'], cell, inp['posfiletype'])
        bas = lattice.fold_basis(cell, bas)
        na = bas.shape[1]
        n1, n2, n3 = inp['ncell']
        natom, mens = na * n1 * n2 * n3, inp['mensemble']
        self.na, self.natom, self.mens = na, natom, mens
        ammom, aemom, landeg = asdio.read_momfile(inp['momfile'], na, inp['landeg_glob'])
        self.anumb = (np.arange(natom, dtype=np.int32) % na) + 1
        self.atype = atype_inp[self.anumb - 1]
        # coordinates, atom order i = i0 + NA*(ix + N1*(iy + N2*iz))  (geometry.f90:440-460)
        idx = np.arange(natom)
        i0, ix, iy, iz = idx % na, (idx // na) % n1, (idx // (na * n1)) % n2, idx // (na * n1 * n2)
        self.coord = (np.outer(cell[0], ix) + np.outer(cell[1], iy) + np.outer(cell[2], iz)) + bas[:, i0]
        if inp['do_prnstruct'] in (1, 2, 4):
            self.out.coord(self.coord, self.atype, self.anumb)
        reduced = inp['do_reduced'] == 'Y'
        nham = na if reduced else natom
        e = host.Engine(device)
        e.set_constants(c['gama'], c['k_bolt'], c['mub'], c['mry'])
        e.set_system(natom, mens, nham, self.anumb if reduced else None)
        tables = [(0, 'exchange', 1, 1, inp['sym'], True), (1, 'dm', 3, 1, 0, False), (2, 'bq', 1, 2, inp['sym'], False)]
        for kind, key, ncomp, lexp, sym, typed in tables:
            if not inp.get(key):
                continue
            if kind == 0 and inp['do_jtensor'] == 1:
                # tensorial exchange: same neighbour map, nine couplings per pair, no neighbour-type filter (kind 3)
                nn, red, xc, nntype = asdio.read_tensorfile(inp[key], atype_inp, bas, cell, inp['maptype'], inp['posfiletype'])
                kind, typed = 3, False
            else:
                nn, red, xc, nntype = asdio.read_pairfile(inp[key], atype_inp, bas, cell, inp['maptype'], inp['posfiletype'], ncomp)
            ns, ca, cs, sh = lattice.stencil(cell, bas, atype_inp, nn, red, sym, nntype if typed else None, ncell=(n1, n2, n3))
            cp = lattice.couplings(ns, ca, sh, atype_inp, xc, ammom, c['mry'], c['mub'], lexp)
            e.build_lattice_table(kind, na, (n1, n2, n3), inp['bc'], ns, ca, cs, cp)
        if inp.get('anisotropy'):
            atyp, an = asdio.read_kfile(inp['anisotropy'], na)
            fc = c['mry'] / c['mub']                    # setup_anisotropies, hamiltonianinit.f90:895-910
            a = self.anumb - 1
            nrm = an[a, 2] ** 2 + an[a, 3] ** 2 + an[a, 4] ** 2
            eaniso = (an[a, 2:5] / np.sqrt(nrm + 1.0e-15)[:, None]).T
            m = ammom[a]
            cub = atyp[a] == 2
            k1 = np.where(cub, fc * an[a, 0] / m ** 4, fc * an[a, 0] / m ** 2)
            k2 = np.where(cub, fc * an[a, 1] / m ** 6, fc * an[a, 1] / m ** 4)
            e.set_anisotropy(atyp[a], np.asfortranarray(eaniso), np.asfortranarray(np.stack([k1, k2])), an[a, 5])
        self.landeg = 0.5 * landeg[self.anumb - 1]       # setup_moment: Landeg = Landeg_ch/2 (magnetizationinit.f90:585)
        self.engine = e
        self._set_field(inp['ip_hfield'] if inp['ip_mode'] != 'N' else inp['hfield'])
        self._llg(inp['sdealgh'], inp['timestep'], inp['damping'], inp['temp'])
        e.commit()
        # ---- moments (magninit, magnetizationinit.f90:141-273) ----
        mmom = np.asfortranarray(np.repeat(ammom[self.anumb - 1][:, None], mens, axis=1))
        if inp['initmag'] == 3:
            emom = np.asfortranarray(np.repeat(aemom[:, self.anumb - 1][:, :, None], mens, axis=2))
        elif inp['initmag'] == 4:
            self.rstep, emom, mmom = asdio.read_restart(inp['restartfile'], natom, mens)
        elif inp['initmag'] == 1:
            # random start: drawn on the host, once, from the reference's own generator seeded with tseed -- the same numbers
            # as the reference, so its Initmag 1 goldens apply to this path unchanged
            e1 = refrng.random_start(na, (n1, n2, n3), inp['tseed'])
            emom = np.asfortranarray(np.repeat(e1[:, :, None], mens, axis=2))
        else:
            raise Unsupported('initmag %d' % inp['initmag'])
        self.mmom0 = mmom.copy(order='F')
        e.set_moments(emom, mmom, self.mmom0)
        self.atype_cell = atype_inp
        if inp['skyno'] == 'T':
            e.set_triangulation(lattice.triangulation(n1, n2, n3, na))          # uppasd.f90:1284-1286
        elif inp['skyno'] == 'Y':
            warnings.warn('skyno Y (finite-difference Pontryagin density) is not on this path; use skyno T (triangulation)')

    def _set_field(self, h):
        f = np.zeros((3, self.natom, self.mens), order='F')
        for a in range(3):
            f[a] = h[a]
        self.hfield = tuple(h)
        self.engine.set_external_field(f)

    def _llg(self, alg, dt, damping, temp):
        if alg not in (1, 5):
            raise Unsupported('SDEalgh %d: this path has 1 (semi-implicit midpoint) and 5 (Depondt)' % alg)
        self.engine.set_llg(alg, dt, landeg=self.landeg, lambda1=damping, temp=temp, mompar=self.inp['mompar'], seed=self.seed)

    # ------------------------------------------------------------------------------------------------
    def run_initial_phase(self):
        inp, e = self.inp, self.engine
        mode = inp['ip_mode']
        if mode == 'N':
            return
        step = 1
        if mode == 'S':
            for nstep, temp, dt, damp in inp['ip_nphase']:
                self._llg(inp['ipsdealgh'], dt, damp, temp)
                e.sd_steps(nstep, first_step=step)
                step += nstep
        else:
            phases = inp['ip_mcanneal'] or [(inp['ip_mcnstep'], inp['ip_temp'])]
            for nsweep, temp in phases:
                e.mc_sweeps(mode, nsweep, temp, first_sweep=step, extfield=inp['ip_hfield'])
                step += nsweep
        self._noise_offset = step
        self._set_field(inp['hfield'])

    # ------------------------------------------------------------------------------------------------
    def _measure(self, mstep, mode):
        """print_averages + print_trajectories of measure() (measurements.f90:106-170)"""
        inp, e = self.inp, self.engine
        msum = None
        if inp['do_avrg'] == 'Y' and (mstep - 1) % inp['avrg_step'] == 0:
            msum = e.measure()
            if inp['do_proj_avrg'] in ('Y', 'A'):
                self.proj_rows += observables.projected_rows(mstep - 1, e.measure_sublattice(self.na), self.natom // self.na,
                                                             self.atype_cell, inp['do_proj_avrg'])
            rows = self.avg.sample(mstep - 1, msum)
            if rows:
                self.out.averages(rows)
                self._flush_proj()
                self._write_restart(mstep, mode)
        if inp['skyno'] == 'T' and (mstep - 1) % inp['skyno_step'] == 0:
            self.sky_rows.append(self.sky.sample(mstep - 1, e.skyrmion_number()))
            if len(self.sky_rows) == inp['skyno_buff']:
                self._flush_sky()
        for t, (atom, tstep, tbuff) in enumerate(inp['trajectories']):
            if (mstep - 1) % tstep == 0:
                v = e.get_atoms([atom])                    # (4, 1, M)
                for k in range(self.mens):
                    self.traj[t][k].append((mstep - 1, v[0, 0, k], v[1, 0, k], v[2, 0, k], v[3, 0, k]))
                if len(self.traj[t][0]) == tbuff:
                    self._flush_traj(t)
        if inp['do_cumu'] == 'Y' and mstep % inp['cumu_step'] == 0:
            if msum is None:
                msum = e.measure()
            row = self.cum.sample(msum, self.last_energy, self.last_exc)
            if row:
                self.out.cumulants(row)
        if inp['do_cumu_proj'] == 'Y' and mstep % inp['cumu_step'] == 0:      # prn_averages.f90:186-189
            rows = self.pcum.sample(e.measure_sublattice(self.na))
            if rows:
                self.out.projcumulants(rows)

    def _energy(self, mstep):
        t = self.engine.energy_terms()                     # (5, M): exc, ani, dm, bq, ext
        tot = t.sum(axis=0)
        self.last_energy, self.last_exc = tot, t[0]
        self.out.totenergy(mstep - 1, dict(tot=tot.mean(), exc=t[0].mean(), ani=t[1].mean(), dm=t[2].mean(), bq=t[3].mean(),
                                           ext=t[4].mean()))

    def _flush_proj(self):
        if self.proj_rows:
            self.out.projavgs(self.proj_rows)
            self.proj_rows = []

    def _flush_sky(self):
        if self.sky_rows:
            self.out.sknumber(self.sky_rows)
            self.sky_rows = []

    def _flush_traj(self, t):
        atom = self.inp['trajectories'][t][0]
        for k in range(self.mens):
            if self.traj[t][k]:
                self.out.trajectory(atom, k + 1, self.traj[t][k])
                self.traj[t][k] = []

    def _write_restart(self, mstep, mode):
        emom, _, mmom = self.engine.get_moments()
        self.out.restart(mstep, mode, emom, mmom)

    def _next_event(self, mstep, last):
        """smallest step > mstep at which measure() or the energy does something (so the steps in between run as one
        batch of kernel launches)"""
        inp = self.inp
        nxt = last + 1
        periods = []
        if inp['do_avrg'] == 'Y' or inp['plotenergy'] > 0:
            periods.append((inp['avrg_step'], 1))          # (m - 1) % p == 0
        for _, tstep, _ in inp['trajectories']:
            periods.append((tstep, 1))
        if inp['skyno'] == 'T':
            periods.append((inp['skyno_step'], 1))
        if inp['do_cumu'] == 'Y' or inp['do_cumu_proj'] == 'Y':
            periods.append((inp['cumu_step'], 0))          # m % p == 0
        for p, off in periods:
            m = ((mstep - off) // p + 1) * p + off
            nxt = min(nxt, m)
        return nxt

    def run_measurement_phase(self):
        inp, e = self.inp, self.engine
        mode = inp['mode']
        self.avg = observables.Averages(self.natom, inp['avrg_buff'])
        self.cum = observables.Cumulants(self.natom, self.mens, inp['temp'], self.c['k_bolt'], self.c['mub'], self.c['mry'],
                                         inp['cumu_buff'], inp['plotenergy'])
        self.traj = [[[] for _ in range(self.mens)] for _ in inp['trajectories']]
        self.proj_rows, self.sky_rows, self.sky = [], [], observables.SkyrmionNumber(self.na)
        self.pcum = observables.ProjectedCumulants(self.atype_cell, self.natom // self.na, self.mens, inp['temp'], self.c['k_bolt'],
                                                   self.c['mub'], inp['cumu_buff'])
        self.last_energy = self.last_exc = None
        off = getattr(self, '_noise_offset', 1) - 1            # keeps the noise counters of the two phases apart
        if mode == 'S':
            self._llg(inp['sdealgh'], inp['timestep'], inp['damping'], inp['temp'])
            mstep, last = self.rstep + 1, self.rstep + inp['nstep']
            if inp['do_bpulse'] in (1, 2, 3, 4):
                # magnetic-field pulse: the whole schedule of the phase goes to the engine once (fields.py, asd_set_time_field)
                P = fields.read_bpulse(inp['bpulsefile'] or os.path.join(self.dir, 'bpulsefile'), inp['do_bpulse'])
                self.bpulse = fields.bpulse_schedule(inp['do_bpulse'], P, inp['timestep'], self.rstep, inp['nstep'])
                e.set_time_field(off + mstep, self.bpulse)
                if inp['plotenergy'] > 0:
                    warnings.warn('do_bpulse: the Zeeman column of totenergy does not include the pulse field')
            while mstep <= last:
                self._measure(mstep, mode)
                if inp['plotenergy'] > 0 and (mstep - 1) % inp['avrg_step'] == 0:
                    self._energy(mstep)
                n = self._next_event(mstep, last) - mstep
                e.sd_steps(n, first_step=off + mstep)
                mstep += n
            self._measure(mstep, mode)                          # sd_driver.f90:839-849: final measure + flush
            if inp['do_bpulse'] in (1, 2, 3, 4):
                e.set_time_field(0, None)
        else:
            mstep, last = 1, inp['mcnstep']
            while mstep <= last:
                self._measure(mstep, mode)
                if inp['plotenergy'] > 0 and (mstep - 1) % inp['avrg_step'] == 0:
                    self._energy(mstep)
                n = self._next_event(mstep, last) - mstep
                e.mc_sweeps(mode, n, inp['temp'], first_sweep=off + mstep, extfield=inp['hfield'])
                mstep += n
        rows = self.avg.flush()
        if rows:
            self.out.averages(rows)
        self._flush_proj()
        self._flush_sky()
        for t in range(len(self.traj)):
            self._flush_traj(t)
        self._write_restart(mstep, mode)
        self.final_step = mstep
        return self

    def run(self):
        self.run_initial_phase()
        return self.run_measurement_phase()

    # ------------------------------------------------------------------------------------------------
    def relax(self, mode='S', nstep=10, temperature=0.0, timestep=1.0e-16, damping=0.5):
        """pyasd.f90:255-298 relax_: nstep steps / sweeps without measurements, returns moments(3,N,M)"""
        e = self.engine
        # the noise / draw counter continues after the phases already run (a fresh counter would replay the initial phase's stream)
        first = getattr(self, '_relax_step', getattr(self, '_noise_offset', 1))
        if mode == 'S':
            self._llg(self.inp['sdealgh'], timestep, damping, temperature)
            e.sd_steps(nstep, first_step=first)
        elif mode in ('M', 'H'):
            e.mc_sweeps(mode, nstep, temperature, first_sweep=first, extfield=self.hfield)
        else:
            raise Unsupported('relax mode %s' % mode)
        self._relax_step = first + nstep
        return e.get_moments()[0]

    def moments(self):
        return self.engine.get_moments()

    def energy(self):
        """total energy per atom in mRy, ensemble mean (pyasd.f90 get_energy_)"""
        return float(self.engine.energy_terms().sum(axis=0).mean())
