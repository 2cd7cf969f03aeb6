"""Random alloys (do_ralloy 1; BASELINE config 3: FeCo random alloy): occupancy and chemistry-dependent couplings.

Host-side product code (numpy).  What the reference does, and where:
  occupancy   setup_chemicaldata, source/System/geometry.f90:190-329 -- for every basis site Ncell uniforms of the host generator
              (re-seeded with tseed just before, uppasd.f90:903-906) rank the cells; species ich takes the next
              nint(conc * Ncell) cells in descending order of the random numbers.  The occupancy of a run is a function of this
              stream, so the reference's own generator (refrng.ReferenceUniform, the MT variant of mtprng.f90) is used.
  couplings   setup_neighbour_hamiltonian, source/Hamiltonian/hamiltonianinit.f90:1075-1084 --
              ncoup(j, i) = xc(atype(i), shell, chem(i), chem(nlist(j, i))) * 2 mRy / mu_B / m(site(i), chem(i)) / m(site(j), chem(j))
              (zero when the product of the two moments is below 1e-6), one row per atom: nHam = Natom.
  moments     setup_moment / magninit (Initmag 3), source/System/magnetizationinit.f90:244-258, 586-597.
Only fully occupied supercells are served (every site's concentrations add up to one): a dilute alloy renumbers the atoms
(acellnumb) and is refused.  The neighbour LIST of a fully occupied alloy is the list of the underlying lattice
(neighbourmap.f90:257-262, 309), so it is built on the device like any other lattice; the chemistry enters the couplings only.
"""
import numpy as np

from . import refrng


def occupancy(na, ncell, nch, chconc, tseed):
    """achtype(Natom): 1-based chemical type of every atom, atom order i0 + NA * cell (geometry.f90:440-460)"""
    ncellt = int(ncell[0]) * int(ncell[1]) * int(ncell[2])
    gen = refrng.ReferenceUniform(tseed)
    ach = np.zeros(na * ncellt, dtype=np.int32)
    for ia in range(na):
        rn = np.array([gen.real2() for _ in range(ncellt)])
        order = np.argsort(-rn, kind='stable')                 # repeated maxloc with the maximum zeroed: descending, first index wins ties
        first = 0
        for ich in range(int(nch[ia])):
            q = int(np.rint(chconc[ia, ich] * ncellt))
            ach[order[first:first + q] * na + ia] = ich + 1
            first += q
    return ach


def mount(nlist, nlistsize, shell, atype, site, chem, xc, ammom, mry, mub):
    """ncoup(z, Natom) for nlist(z, Natom) (1-based, rows beyond nlistsize ignored): shell(z, Natom) = 0-based shell of every entry,
    atype / site / chem (Natom) 1-based type, basis site and species, xc(NT, maxshell, Nch, Nch), ammom(NA, Nch)."""
    z, n = nlist.shape
    nb = np.clip(nlist, 1, n) - 1
    mi = ammom[site - 1, chem - 1][None, :]
    mj = ammom[site[nb] - 1, chem[nb] - 1]
    x = xc[atype[None, :] - 1, shell, chem[None, :] - 1, chem[nb] - 1]
    fc2 = 2.0 * mry / mub
    with np.errstate(divide='ignore', invalid='ignore'):
        c = x * fc2 / mi / mj                                   # left to right, as the Fortran expression
    c = np.where(np.abs(mi * mj) < np.float64(np.float32(1e-6)), 0.0, c)
    c = np.where(np.arange(z)[:, None] < nlistsize[None, :], c, 0.0)
    return np.asfortranarray(c)
