"""BASELINE.json config 3 as a parity case: a 2-D triangular lattice with nearest-neighbour Heisenberg exchange, interfacial
Dzyaloshinskii-Moriya vectors, uniaxial anisotropy and a field along -z (the ingredients of
examples/SpecialFeatures/SkyrmionLattice), set up FROM FILES by the product driver (readers, stencil, on-device tables, 2-D
bricks, open boundary along z) and compared with the oracle built from the same files: tables bit-exact, field and T = 0
LLG relaxation to 1e-12, heat-bath annealing on observables."""
import os

import numpy as np
import pytest

from oracle import inputs as oinputs
from oracle import orc

pytestmark = pytest.mark.gpu

A1, A2 = (1.0, 0.0, 0.0), (-0.5, 0.8660254037844386, 0.0)
NB = [(1, 0), (0, 1), (-1, -1), (-1, 0), (0, -1), (1, 1)]     # the six nearest neighbours in units of (A1, A2)


def _write(d, ncell=(64, 32, 1), temp=0.0, mode='S', nstep=300, extra=''):
    vec = lambda n: (n[0] * A1[0] + n[1] * A2[0], n[0] * A1[1] + n[1] * A2[1], 0.0)
    with open(os.path.join(d, 'posfile'), 'w') as fh:
        fh.write('1 1 0.0 0.0 0.0\n')
    with open(os.path.join(d, 'momfile'), 'w') as fh:
        fh.write('1 1 1.5 0.1 0.05 1.0\n')
    with open(os.path.join(d, 'jfile'), 'w') as fh:
        for n in NB:
            fh.write('1 1 %.10f %.10f %.10f 1.0\n' % vec(n))
    with open(os.path.join(d, 'dmfile'), 'w') as fh:
        for n in NB:
            r = np.array(vec(n))
            dm = 0.35 * np.cross([0.0, 0.0, 1.0], r / np.linalg.norm(r))     # interfacial DMI: D perpendicular to bond and z
            fh.write('1 1 %.10f %.10f %.10f %.10f %.10f %.10f\n' % (*r, *dm))
    with open(os.path.join(d, 'kfile'), 'w') as fh:
        fh.write('1 1 0.05 0.0 0.0 0.0 1.0 0.0\n')
    with open(os.path.join(d, 'inpsd.dat'), 'w') as fh:
        fh.write('''simid skyrm_2D
ncell %d %d %d
BC P P 0
cell %.10f %.10f %.10f
     %.10f %.10f %.10f
     0.0 0.0 1.0
Sym 0
posfile ./posfile
momfile ./momfile
exchange ./jfile
dm ./dmfile
anisotropy ./kfile
do_reduced Y
Mensemble 2
Initmag 3
SDEalgh 1
mode %s
temp %g
hfield 0.0 0.0 -2.5
damping 0.3
timestep 1.0d-16
Nstep %d
mcNstep %d
do_avrg Y
avrg_step 50
plotenergy 1
%s
''' % (*ncell, *A1, *A2, mode, temp, nstep, nstep, extra))
    return os.path.join(d, 'inpsd.dat')


def _oracle_system(path, mens):
    inp = oinputs.read_inpsd(path)
    inp['mensemble'] = mens
    bas, atype = oinputs.read_positions(inp['files']['posfile'], inp['cell'], inp['posfiletype'])
    am, ae, lg = oinputs.read_moments(inp['files']['momfile'], 1, inp['landeg_glob'])
    pair = lambda key, nc, typed: (lambda S: oinputs.read_pair_file(inp['files'][key], 1, atype, S['bas'], inp['cell'], inp['maptype'],
                                                                    inp['posfiletype'], nc, typed))
    aniso = oinputs.read_anisotropy(inp['files']['anisotropy'], 1)
    return inp, orc.build_system(inp, bas, atype, am, ae, lg, pair('exchange', 1, True), dm=pair('dm', 3, False), aniso=aniso)


def test_triangular_dmi_field_relaxation(tmp_path):
    from uppasd_b200 import driver
    path = _write(str(tmp_path))
    sim = driver.Simulation(path)
    inp, S = _oracle_system(path, 2)
    e = sim.engine
    # 64 x 32 cells, one basis atom: 32 x 8 bricks, four of them (adjacent in y) per 1024-slot tile -> run kernel with the DM
    # neighbours in the shared-memory gather list
    info = e.layout_info()
    assert info['runs'] == 4 and info['tile_slots'] == 1024 and info['extra_staged'] == 1, info
    # tables built on the device from the files = the oracle's, bit for bit (2-D bricks, open boundary along z)
    for kind, key in ((0, 'exchange'), (1, 'dm')):
        lst, size, coup = e.get_table(kind)
        assert np.array_equal(lst, S[key]['list']) and np.array_equal(size, S[key]['listsize']) and np.array_equal(coup, S[key]['coup'])
    # a non-trivial start: the driver's uniform state, twisted site by site
    N = S['Natom']
    ph = 2 * np.pi * np.modf(np.arange(1, N + 1) * 0.6180339887)[0]
    e0 = np.stack([0.6 * np.cos(ph), 0.6 * np.sin(ph), np.full(N, 0.8)])[:, :, None].repeat(2, axis=2)
    e0[:, :, 1] = e0[::-1, :, 1]
    e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    e.set_moments(S['emom'], S['mmom'])
    beff, en = e.effective_field()
    rb, ren = orc.effective_field(S)
    assert np.abs(beff - rb).max() <= 1e-12 * np.abs(rb).max()
    terms = e.energy_terms()
    assert abs(terms.sum(axis=0).mean() - ren / N / 2) <= 1e-11 * abs(ren / N / 2) + 1e-12    # ren sums both ensembles
    assert abs(terms[2]).max() > 1e-3 and abs(terms[1]).max() > 1e-4 and abs(terms[4]).max() > 1e-2    # DM, anisotropy, Zeeman all act
    st = orc.SdState(S, 1, inp['timestep'], inp['damping'])
    e.set_llg(1, inp['timestep'], landeg=S['Landeg'], lambda1=inp['damping'], temp=0.0)
    e.sd_steps(300)
    for _ in range(300):
        st.step()
    assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12
    e_after = e.energy_terms().sum(axis=0)
    assert (e_after < terms.sum(axis=0)).all()


def test_triangular_heat_bath_anneal_observables(tmp_path):
    """heat-bath sweeps at 30 K from the polarised start, GPU (colour-parallel) vs oracle (random sequential order)"""
    from uppasd_b200 import driver
    path = _write(str(tmp_path), ncell=(32, 16, 1))
    sim = driver.Simulation(path)
    inp, S = _oracle_system(path, 2)
    e = sim.engine
    T = 30.0
    e.mc_sweeps('H', 150, T)
    ge, gm = [], []
    for r in range(40):
        e.mc_sweeps('H', 5, T, first_sweep=151 + 5 * r)
        m, en = e.measure(energy=True)
        ge.append(en / S['Natom']); gm.append(m[2] / S['Natom'])
    ge, gm = np.array(ge), np.array(gm)
    rm, re_, (emom, emomM, mmom) = orc.mc_run(S, 'H', T, 350, seed=4, sample_every=5, burn=150)
    assert abs(ge.mean() - re_.mean()) < 0.03 * abs(re_.mean()) + 5 * ge.std() / np.sqrt(ge.size)
    assert abs(gm.mean() - emomM[2].mean()) < 0.05 * abs(emomM[2].mean()) + 0.02


def test_skyrmion_number_and_sublattice_sums(tmp_path):
    """asd_skyrmion_number against the oracle's pontryagin_tri on the triangulation delaunay_tri_tri builds (topology.f90),
    for a Neel skyrmion written into the triangular lattice (Q = -1), a twisted texture, both device layouts (LLG tiles and
    Monte Carlo colour order), and the sknumber.*.out the driver writes with skyno T; asd_measure_sublattice against numpy."""
    from uppasd_b200 import asdio, driver, lattice
    n1, n2 = 64, 32
    path = _write(str(tmp_path), ncell=(n1, n2, 1), nstep=100, extra='skyno T\nskyno_step 50\nskyno_buff 2\ndo_proj_avrg A')
    sim = driver.Simulation(path)
    e = sim.engine
    N = n1 * n2
    simp = orc.delaunay_tri_tri(n1, n2, 1, 1)
    assert np.array_equal(simp, lattice.triangulation(n1, n2, 1, 1))
    idx = np.arange(N)
    pos = np.outer(idx % n1, A1[:2]) + np.outer(idx // n1, A2[:2])
    d = pos - pos.mean(axis=0)
    r, phi = np.hypot(d[:, 0], d[:, 1]), np.arctan2(d[:, 1], d[:, 0])
    th = np.pi * np.exp(-r / 5.0)
    sk = np.stack([np.sin(th) * np.cos(phi), np.sin(th) * np.sin(phi), np.cos(th)])
    ph = 2 * np.pi * np.modf(np.arange(1, N + 1) * 0.6180339887)[0]
    tw = np.stack([0.6 * np.cos(ph), 0.6 * np.sin(ph), np.full(N, 0.8)])
    tw /= np.sqrt((tw ** 2).sum(axis=0))
    emom = np.asfortranarray(np.stack([sk, tw], axis=2))
    mmom = np.full((N, 2), 1.5, order='F')
    e.set_moments(emom, mmom)
    q = e.skyrmion_number()
    qref, per = orc.pontryagin_tri(emom, simp)
    assert abs(per[0] + 1.0) < 1e-10
    assert np.abs(q - per).max() <= 1e-12 * max(1.0, np.abs(per).max()), (q, per)
    ms = e.measure_sublattice(1)
    assert np.abs(ms[:, 0, :] - (emom * mmom[None]).sum(axis=1)).max() <= 1e-9
    # the same state in the Monte Carlo (colour-major) layout: zero sweeps move the state there
    e.mc_sweeps('H', 0, 1.0)
    e.mc_sweeps('M', 1, 1e-9)
    q2 = e.skyrmion_number()
    em2 = e.get_moments()[0]
    assert np.abs(q2 - orc.pontryagin_tri(em2, simp)[1]).max() <= 1e-12 * max(1.0, np.abs(q2).max())
    # driver output: sknumber.<simid>.out rows at 0, 50, 100 with the running mean of prn_skyno
    e.set_moments(emom, mmom)
    sim.run()
    rows = asdio.read_out(os.path.join(str(tmp_path), 'sknumber.skyrm_2D.out'))
    assert [int(x[0]) for x in rows] == [0, 50, 100]
    assert abs(rows[0][1] - qref) <= 1e-8                      # printed with f16.8; NA = 1
    assert abs(rows[2][2] - np.mean([x[1] for x in rows])) <= 2e-8
    pr = [x for x in asdio.read_out(os.path.join(str(tmp_path), 'projavgs.skyrm_2D.out')) if int(x[0]) == 0]
    av = _first_averages(str(tmp_path))
    assert len(pr) == 1 and abs(pr[0][2] - av[4]) <= 1e-8 and abs(pr[0][6] - av[3]) <= 1e-8


def _first_averages(d):
    from uppasd_b200 import asdio
    return asdio.read_out(os.path.join(d, 'averages.skyrm_2D.out'))[0]
