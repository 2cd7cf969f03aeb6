"""Observables of the measurement phases, accumulated on the host from the per-ensemble sums the device returns
(asd_measure / asd_energy_terms) -- SURVEY 8 f-2.  The estimators are the reference's, not textbook ones:

  averages   buffer_avrg / prn_avrg      source/Measurement/prn_averages.f90:414-456, 569-656
  cumulants  calc_and_print_cumulant     source/Measurement/prn_averages.f90:919-1095 -- WEIGHTED running means with a
             linearly growing weight (cumuw += 1 per sample per ensemble), U = 1 - <m^4>/(3 <m^2>^2),
             chi = (<m^2> - <m>^2) mu_B^2 N / (k_B^2 T), C_v from the energy variance in mRy
  projavgs   buffer_proj_avrg / prn_proj_avrg  source/Measurement/prn_averages.f90:462-512, 662-802
  sknumber   buffer_skyno_tri / prn_skyno      source/Measurement/prn_topology.f90:660-700, 296-345
  projcumulants  calc_and_print_cumulant_proj  source/Measurement/prn_averages.f90:1239-1333 -- PLAIN running means per atom
             type (unlike the weighted means of the total cumulants)
"""
import numpy as np


class Averages:
    """Buffered <M> rows of averages.<simid>.out."""

    def __init__(self, natom, buff=10):
        self.natom, self.buff, self.rows = natom, buff, []

    def sample(self, it, msum):
        """msum(3, M) = sum_i emomM(:, i, k); returns the rows to print when the buffer is full, else None"""
        av = np.asarray(msum, dtype=np.float64) / self.natom          # (3, M)
        m = np.sqrt((av ** 2).sum(axis=0))                           # |<M>| per ensemble
        mm = m.mean()
        var = (m ** 2).mean() - mm ** 2
        sd = 0.0 if var < 1.0e-14 else float(np.sqrt(var))           # dbl_tolerance filter (prn_averages.f90:632-636)
        self.rows.append((it, av[0].mean(), av[1].mean(), av[2].mean(), mm, sd))
        if len(self.rows) == self.buff:
            return self.flush()
        return None

    def flush(self):
        out, self.rows = self.rows, []
        return out


class Cumulants:
    """Binder cumulant, susceptibility, specific heat with the reference's weighted running means."""

    def __init__(self, natom, mensemble, temp, k_bolt, mub, mry, buff=10, plotenergy=0):
        self.n, self.m, self.temp, self.kb, self.mub, self.mry = natom, mensemble, temp, k_bolt, mub, mry
        self.buff, self.plotenergy = buff, plotenergy
        self.cumuw = self.cumutotw = 0.0
        self.navrg = 0
        self.avm = self.avm2 = self.avm4 = 0.0
        self.ave = self.ave2 = self.avexc = 0.0
        self.binder = self.chi = self.cv = 0.0

    def sample(self, msum, energy=None, exc=None):
        """msum(3, M); energy / exc: per-ensemble total / exchange energy per atom in mRy (the values of the LAST
        energy evaluation, as in the reference where calc_energy runs on its own cadence).  Returns the row to print
        (count, <M>, <M^2>, <M^4>, U, chi, Cv, <E>, <E_exc>, <E_lsf>) or None."""
        msum = np.asarray(msum, dtype=np.float64)
        avm = avm2 = avm4 = ave = ave2 = avexc = 0.0
        for k in range(self.m):
            me = float(np.sqrt((msum[:, k] ** 2).sum())) / self.n
            m2 = me * me
            m4 = m2 * m2
            self.cumuw += 1.0
            w, tw = self.cumuw, self.cumutotw
            avm = (self.avm * tw + me * w) / (tw + w)
            avm2 = (self.avm2 * tw + m2 * w) / (tw + w)
            avm4 = (self.avm4 * tw + m4 * w) / (tw + w)
            self.binder = 1.0 - (avm4 / 3.0 / avm2 ** 2)
            self.avm, self.avm2, self.avm4 = avm, avm2, avm4
            if self.temp > 0.0:
                self.chi = (avm2 - avm ** 2) * self.mub ** 2 * self.n / (self.kb ** 2) / self.temp
            else:
                self.chi = (avm2 - avm ** 2) * self.n * self.mub ** 2 / self.kb
            if self.plotenergy > 0 and energy is not None:
                ek = float(energy[k])
                ave = (self.ave * tw + ek * w) / (tw + w)
                ave2 = (self.ave2 * tw + ek * ek * w) / (tw + w)
                avexc = (self.avexc * tw + (float(exc[k]) if exc is not None else 0.0) * w) / (tw + w)
                self.cv = ((ave2 - ave ** 2) * self.mry ** 2 * self.n / (self.kb ** 2) / self.temp ** 2) if self.temp > 0.0 else 0.0
                self.ave, self.ave2, self.avexc = ave, ave2, avexc
            else:
                self.cv = 0.0
            self.navrg += 1
            self.cumutotw += w
        count = self.navrg // self.m
        if (count - 1) % self.buff == 0:
            return (count, self.avm, self.avm2, self.avm4, self.binder, self.chi, self.cv, self.ave, self.avexc, 0.0)
        return None


def projected_rows(it, msum_na, ncells, atype_cell, mode='Y'):
    """Rows of projavgs.<simid>.out for one sample: msum_na(3, NA, M) = sums of emomM per basis atom.  mode 'Y': one row per
    atom TYPE, the sums of the basis atoms of that type divided by the number of CELLS (not by the number of atoms of the
    type -- prn_averages.f90:713-722); mode 'A': one row per basis atom.  Row = (iter, proj, <M>, M_stdv, <M>_x, <M>_y, <M>_z)."""
    msum_na = np.asarray(msum_na, dtype=np.float64)
    na, mens = msum_na.shape[1], msum_na.shape[2]
    atype_cell = np.asarray(atype_cell)
    if mode == 'Y':
        nproj = int(atype_cell.max())
        v = np.zeros((3, nproj, mens))
        for i_na in range(na):
            v[:, atype_cell[i_na] - 1, :] += msum_na[:, i_na, :] / ncells
    else:
        nproj = na
        v = msum_na / ncells
    rows = []
    for k in range(nproj):
        m = np.sqrt((v[:, k, :] ** 2).sum(axis=0))             # per ensemble
        mm = m.mean()
        var = (m ** 2).mean() - mm ** 2
        rows.append((it, k + 1, mm, 0.0 if var < 0 else float(np.sqrt(var)), v[0, k].mean(), v[1, k].mean(), v[2, k].mean()))
    return rows


class SkyrmionNumber:
    """Running mean / variance of the skyrmion number as prn_skyno accumulates them (Welford update, prn_topology.f90:321-327)."""

    def __init__(self, na):
        self.na, self.count, self.avg, self.var = na, 0, 0.0, 0.0

    def sample(self, it, q):
        """q: per-ensemble sums of solid angles / 4 pi (asd_skyrmion_number); the printed number is their ensemble mean / NA
        (pontryagin_tri divides by Mensemble, buffer_skyno_tri by NA).  Returns the row (iter, Skx num, Skx avg, Skx std)."""
        x = float(np.mean(q)) / self.na
        self.count += 1
        prev = self.avg
        self.avg = prev + (x - prev) / self.count
        self.var += (x - prev) * (x - self.avg)
        return (it, x, self.avg, self.var / self.count)


class ProjectedCumulants:
    """Per-type Binder cumulant and susceptibility (do_cumu_proj Y): plain running means over all samples and ensembles."""

    def __init__(self, atype_cell, ncells, mensemble, temp, k_bolt, mub, buff=10):
        self.atype = np.asarray(atype_cell)
        self.nt = int(self.atype.max())
        self.ncount = np.array([int((self.atype == it + 1).sum()) * ncells for it in range(self.nt)], dtype=np.float64)
        self.m, self.temp, self.kb, self.mub, self.buff = mensemble, temp, k_bolt, mub, buff
        self.n = 0
        self.c1 = np.zeros(self.nt)
        self.c2 = np.zeros(self.nt)
        self.c4 = np.zeros(self.nt)

    def sample(self, msum_na):
        """msum_na(3, NA, M): sums of emomM per basis atom (asd_measure_sublattice).  Returns the rows to print
        (count, type, <M>, <M^2>, <M^4>, U, chi) when mod(count - 1, cumu_buff) == 0, else None."""
        msum_na = np.asarray(msum_na, dtype=np.float64)
        u = np.zeros(self.nt)
        chi = np.zeros(self.nt)
        for k in range(self.m):
            for it in range(self.nt):
                v = msum_na[:, self.atype == it + 1, k].sum(axis=1)
                me = float(np.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2)) / self.ncount[it]
                m2 = me * me
                m4 = m2 * m2
                t1 = (self.n * self.c1[it] + me) / (self.n + 1)
                t2 = (self.n * self.c2[it] + m2) / (self.n + 1)
                t4 = (self.n * self.c4[it] + m4) / (self.n + 1)
                u[it] = 1.0 - (t4 / 3.0 / t2 ** 2)
                self.c1[it], self.c2[it], self.c4[it] = t1, t2, t4
                if self.temp > 0.0:
                    chi[it] = (t2 - t1 ** 2) * self.mub ** 2 * self.ncount[it] / (self.kb ** 2) / self.temp
                else:
                    chi[it] = (t2 - t1 ** 2) * self.ncount[it] * self.mub ** 2 / self.kb
            self.n += 1
        count = self.n // self.m
        if (count - 1) % self.buff == 0:
            return [(count, it + 1, self.c1[it], self.c2[it], self.c4[it], u[it], chi[it]) for it in range(self.nt)]
        return None
