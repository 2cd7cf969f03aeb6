"""On-device table construction must be bit-identical to the reference's setup_nm + mount (via the oracle)."""
import numpy as np
import pytest

from oracle import inputs, orc
from util import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _stage_launch_path(monkeypatch):
    """this module is about the kernels of LARGE systems, built small on purpose: keep the resident small-system kernel out"""
    monkeypatch.setenv('ASD_RESIDENT', '0')


def _build(fx, over):
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], **over)
    return args, orc.build_system(*args)


CASES = [
    ('kagome', dict()),
    ('kagome', dict(ncell=(5, 7, 1), do_reduced='N', bc=('0', 'P', '0'))),
    ('megatest', dict()),
    ('megatest', dict(ncell=(2, 3, 5), do_reduced='N')),
    ('megatest', dict(ncell=(3, 3, 3), do_reduced='N', bc=('0', '0', '0'))),
    ('feco', dict()),
    ('feco', dict(ncell=(3, 4, 2), do_reduced='N')),           # tiny periodic box: duplicate neighbours dropped
    ('bccfe_cuda', dict(ncell=(12, 10, 8))),
    ('bccfe_cuda', dict(ncell=(4, 4, 4), do_reduced='N', bc=('P', '0', 'P'))),
    ('cluster', dict()),                                       # ONE cell, 43 basis atoms, BC 0 0 0: hops forced to zero; BQ table (lexp 2)
    ('heisstripe', dict()),                                    # BC 0 0 P stripe
    ('heischainaf', dict()),                                   # two atom types
    ('scsurf', dict()),                                        # atomic units, maptype 2, DM
    ('kagome_cuda', dict(ncell=(12, 9, 1))),                   # tensorial exchange (nine couplings per pair), do_reduced N
    ('kagome_cuda', dict(ncell=(7, 5, 1), do_reduced='Y')),
]


@pytest.mark.parametrize('name,over', CASES)
def test_device_tables_bit_exact(name, over):
    from uppasd_b200 import host, lattice
    fx, _, _ = load_golden(name)
    args, S = _build(fx, over)
    inp = args[0]
    kinds = [(0, 'exchange', args[6], 1, 1, inp['sym'], True)]
    if inp.get('do_jtensor', 0) == 1:
        kinds = [(3, 'exchange', args[6], 9, 1, inp['sym'], False)]   # builder kind 3 fills the exchange table with j_tens
    if args[7] is not None:
        kinds.append((1, 'dm', args[7], 3, 1, 0, False))
    if args[8] is not None:
        kinds.append((2, 'bq', args[8], 1, 2, inp['sym'], False))     # hamiltonianinit.f90:485-487: BQ map WITH the lattice symmetry
    c = orc.consts(S)
    e = host.Engine()
    e.set_system(S['Natom'], 1, S['nHam'], S['aHam'])
    for kind, key, mk, ncomp, lexp, sym, typed in kinds:
        nn, red, xc, nntype = mk(S)
        ns, ca, cs, sh = lattice.stencil(inp['cell'], S['bas'], S['atype_inp'], nn, red, sym, nntype if typed else None, ncell=inp['ncell'])
        cp = lattice.couplings(ns, ca, sh, S['atype_inp'], xc, S['ammom_inp'], c['mry'], c['mub'], lexp)
        e.build_lattice_table(kind, S['NA'], inp['ncell'], inp['bc'], ns, ca, cs, cp)
        lst, size, coup = e.get_table(0 if kind == 3 else kind)
        ref = S[key]
        assert lst.shape == ref['list'].shape, (lst.shape, ref['list'].shape)
        assert np.array_equal(size, ref['listsize'])
        assert np.array_equal(lst, ref['list'])
        assert np.array_equal(coup, ref['coup'])      # bit-exact couplings


def test_device_tables_drive_same_dynamics():
    from uppasd_b200 import host, lattice
    fx, inp, S = load_golden('bccfe_cuda')
    args = inputs.load_fixture(fx)
    nn, red, xc, nntype = args[6](S)
    ns, ca, cs, sh = lattice.stencil(inp['cell'], S['bas'], S['atype_inp'], nn, red, inp['sym'], nntype)
    cp = lattice.couplings(ns, ca, sh, S['atype_inp'], xc, S['ammom_inp'], orc.CONST['mry'], orc.CONST['mub'])
    e = host.Engine()
    e.set_constants(*(orc.CONST[k] for k in ('gama', 'k_bolt', 'mub', 'mry')))
    e.set_system(S['Natom'], 1, S['nHam'], S['aHam'])
    e.build_lattice_table(0, S['NA'], inp['ncell'], inp['bc'], ns, ca, cs, cp)
    e.set_llg(1, inp['timestep'], landeg=S['Landeg'], lambda1=inp['damping'], temp=0.0)
    e.set_moments(S['emom'], S['mmom'])
    e.commit()
    e.sd_steps(100)
    st = orc.SdState(S, 1, inp['timestep'], inp['damping'])
    for _ in range(100):
        st.step()
    emom, _, _ = e.get_moments()
    assert np.abs(emom - st.emom).max() <= 1e-12


def _bcc_engine(S, inp, args, solver, temp, seed=11):
    from uppasd_b200 import host, lattice
    nn, red, xc, nntype = args[6](S)
    ns, ca, cs, sh = lattice.stencil(inp['cell'], S['bas'], S['atype_inp'], nn, red, inp['sym'], nntype)
    cp = lattice.couplings(ns, ca, sh, S['atype_inp'], xc, S['ammom_inp'], orc.CONST['mry'], orc.CONST['mub'])
    e = host.Engine()
    e.set_constants(*(orc.CONST[k] for k in ('gama', 'k_bolt', 'mub', 'mry')))
    e.set_system(S['Natom'], S['emom'].shape[2], S['nHam'], S['aHam'])
    e.build_lattice_table(0, S['NA'], inp['ncell'], inp['bc'], ns, ca, cs, cp)
    e.set_llg(solver, inp['timestep'], landeg=S['Landeg'], lambda1=0.3, temp=temp, seed=seed)
    e.set_moments(S['emom'], S['mmom'])
    e.commit()
    return e


@pytest.mark.parametrize('ncell,tile', [((64, 4, 4), 1024), ((64, 6, 4), 1024), ((128, 8, 2), 1024), ((64, 6, 8), 1024), ((96, 8, 6), 1024), ((40, 8, 8), 1024)])
def test_run_kernel_against_oracle_and_staged_kernel(ncell, tile, monkeypatch):
    """The run-compressed register-blocked kernel (asd_runs.cuh) on lattices whose x extent is a multiple of 32 -- one,
    several and partly empty super-bricks: field-level parity is implied by the trajectory (every step evaluates the
    field twice); both solvers against the oracle at T = 0 to 1e-12, and against the one-atom-per-thread staged
    kernel with the same noise stream at 300 K.  The last three shapes take the moment-plane instantiation (gather list staged from
    emomM[M][3][Npad] with cp.async, padded union rows) on partly empty super-bricks; (40, 8, 8) has a partly filled brick along x (8 of 32 cells)."""
    fx, _, _ = load_golden('bccfe_cuda')
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], ncell=ncell, mensemble=2, do_reduced='Y')
    S = orc.build_system(*args)
    inp = args[0]
    rng = np.random.default_rng(3)
    e0 = rng.normal(size=(3, S['Natom'], 2)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    for solver in (1, 5):
        monkeypatch.delenv('ASD_RUNS', raising=False)
        e = _bcc_engine(S, inp, args, solver, 0.0)
        info = e.layout_info()
        assert info['runs'] == 4 and info['tile_slots'] == tile and info['union'] <= 200, info
        assert info['planes'] == 1 or ncell[2] < 6, info     # the last two shapes are there for the moment-plane instantiation
        beff, _ = e.effective_field()
        rb, _ = orc.effective_field(S)
        assert np.abs(beff - rb).max() <= 1e-12 * np.abs(rb).max()
        st = orc.SdState(S, solver, inp['timestep'], 0.3)
        e.sd_steps(40)
        for _ in range(40):
            st.step()
        assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12, (solver, ncell)
        # thermal: same Philox stream in both kernels (the run kernel draws it while the gather list is in flight)
        er = _bcc_engine(S, inp, args, solver, 300.0)
        monkeypatch.setenv('ASD_RUNS', '0')
        es = _bcc_engine(S, inp, args, solver, 300.0)
        assert es.layout_info()['runs'] == 0 and es.layout_info()['staged'] == 1
        er.sd_steps(25)
        es.sd_steps(25)
        a, b = er.get_moments()[0], es.get_moments()[0]
        assert np.abs(a - b).max() <= 1e-11, (solver, ncell, np.abs(a - b).max())
        assert np.abs(a - e0).max() > 1e-3          # the state did move
        assert np.allclose(er.measure(), es.measure(), rtol=1e-10, atol=1e-8)   # fused per-tile moment sums


def test_irregular_lattice_keeps_the_staged_kernel(monkeypatch):
    """x extent not a multiple of 32 (or a single periodic 32-cell run, whose wrap splits the warp's neighbour run): the
    regularity check of the run table fails on the device and the layout keeps the staged kernel -- same results as the
    oracle."""
    fx, _, _ = load_golden('bccfe_cuda')
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], ncell=(12, 6, 4), do_reduced='Y')
    S = orc.build_system(*args)
    monkeypatch.setenv('ASD_RUNS', '256')          # ask for the run kernel: the device-side check must refuse it
    e = _bcc_engine(S, args[0], args, 1, 0.0)
    info = e.layout_info()
    assert info['runs'] == 0 and info['staged'] == 1 and info['tile_slots'] == 256, info
    st = orc.SdState(S, 1, args[0]['timestep'], 0.3)
    e.sd_steps(30)
    for _ in range(30):
        st.step()
    assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12


@pytest.mark.parametrize('reduced', ['Y', 'N'])
def test_host_tables_with_lattice_hint_take_the_fast_path(reduced):
    """The drop-in case: nlist / ncoup built by the (Fortran) host in the reference's atom order, plus the supercell shape
    (asd_set_lattice_hint / fortrandata_setlattice_).  The engine stores the atoms in bricks; with do_reduced Y the
    device-side check accepts the host's table for the run kernel, with do_reduced N the staged kernel runs on the
    brick order.  Same trajectory as the oracle and as the same engine without the hint."""
    from uppasd_b200 import host
    fx, _, _ = load_golden('bccfe_cuda')
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], ncell=(64, 4, 6), mensemble=2, do_reduced=reduced, hfield=(0.1, 0.0, 0.2))
    S = orc.build_system(*args)
    inp = args[0]
    rng = np.random.default_rng(5)
    e0 = rng.normal(size=(3, S['Natom'], 2)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    hint = (S['NA'], inp['ncell'], inp['bc'])
    for solver in (1, 5):
        eh = host.engine_from_system(S, orc.CONST, sdealgh=solver, delta_t=inp['timestep'], damping=0.2, lattice_hint=hint)
        ep = host.engine_from_system(S, orc.CONST, sdealgh=solver, delta_t=inp['timestep'], damping=0.2)
        ih, ip = eh.layout_info(), ep.layout_info()
        assert ip['runs'] == 0 and ip['tile_slots'] == 256
        if reduced == 'Y':
            assert ih['runs'] == 4 and ih['tile_slots'] == 1024, ih
        else:
            assert ih['runs'] == 0 and ih['staged'] == 1, ih
        rb, _ = orc.effective_field(S)
        for e in (eh, ep):
            beff, _ = e.effective_field()
            assert np.abs(beff - rb).max() <= 1e-12 * np.abs(rb).max()
        st = orc.SdState(S, solver, inp['timestep'], 0.2)
        for _ in range(30):
            st.step()
        for e in (eh, ep):
            e.sd_steps(30)
            assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12, (solver, reduced)
        # a shape that does not match the tables is ignored, not trusted
        ew = host.engine_from_system(S, orc.CONST, sdealgh=solver, delta_t=inp['timestep'], damping=0.2,
                                     lattice_hint=(S['NA'], (64, 4, 5), inp['bc']))
        assert ew.layout_info()['runs'] == 0
        ew.sd_steps(30)
        assert np.abs(ew.get_moments()[0] - st.emom).max() <= 1e-12


def test_legacy_boundary_with_lattice_shape():
    """fortrandata_setlattice_ + the reference's own call sequence: the Fortran-built table of a bcc supercell runs on the
    run-compressed kernel behind the legacy symbols; final state = oracle."""
    from uppasd_b200 import host
    fx, _, _ = load_golden('bccfe_cuda')
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], ncell=(64, 4, 4), do_reduced='Y')
    S = orc.build_system(*args)
    inp = args[0]
    rng = np.random.default_rng(9)
    e0 = rng.normal(size=(3, S['Natom'], 1)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    fh = host.FortranHost(S, orc.CONST, sdealgh=1, nstep=60, delta_t=inp['timestep'], damping=0.3, avrg_step=20,
                          lattice=(S['NA'], inp['ncell'], inp['bc'])).run()
    info = fh.layout_info()
    assert info['runs'] == 4 and info['tile_slots'] == 1024, info
    st = orc.SdState(S, 1, inp['timestep'], 0.3)
    for _ in range(60):
        st.step()
    assert np.abs(fh.arr['emom'] - st.emom).max() <= 1e-12
    assert sorted(fh.averages) == [0, 20, 40, 60]


def test_fixed_moment_list_in_the_run_kernel():
    """red_atom_list on a device-built lattice whose layout carries the run-compressed table: the engine routes a fixed-moment
    run to the direct one-atom-per-thread kernel (the only large-system kernel compiled with the frozen mask); frozen atoms bit
    for bit, the rest against the oracle to 1e-12, both solvers; clearing the list returns to the run kernel."""
    fx, _, _ = load_golden('bccfe_cuda')
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], ncell=(64, 4, 4), mensemble=2, do_reduced='Y')
    S = orc.build_system(*args)
    inp = args[0]
    rng = np.random.default_rng(5)
    e0 = rng.normal(size=(3, S['Natom'], 2)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    N = S['Natom']
    fro = (np.arange(N) % 7 == 3)
    red = np.arange(1, N + 1)[~fro]
    for solver in (1, 5):
        e = _bcc_engine(S, inp, args, solver, 0.0)
        assert e.layout_info()['runs'] == 4
        e.set_evolving_atoms(red)
        st = orc.SdState(S, solver, inp['timestep'], 0.3, red_atom_list=red)
        e.sd_steps(30)
        for _ in range(30):
            st.step()
        emom = e.get_moments()[0]
        assert np.array_equal(emom[:, fro, :], S['emom'][:, fro, :])
        assert np.abs(emom - st.emom).max() <= 1e-12, solver
        # the per-tile moment sums are not fused on this path: asd_measure still returns the right sums
        assert np.allclose(e.measure(), e.get_moments()[1].sum(axis=1), rtol=1e-12, atol=1e-9)
        e.set_evolving_atoms(None)
        e.set_moments(S['emom'], S['mmom'])
        st = orc.SdState(S, solver, inp['timestep'], 0.3)
        e.sd_steps(10)
        for _ in range(10):
            st.step()
        assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12, solver


@pytest.mark.parametrize('solver', [1, 5])
def test_headline_launch_geometry_64_against_oracle(solver, monkeypatch):
    """bcc Fe 64^3 (524 288 spins, 512 tiles of 1024 slots: several waves of CTAs, L2 look-ahead one wave ahead, dependent
    launches) -- the launch geometry of the headline benchmark, T = 0, 10 steps, both solvers, against the oracle to 1e-12."""
    import bench
    monkeypatch.delenv('ASD_RESIDENT', raising=False)
    nc = (64, 64, 64)
    S = bench.oracle_bcc(nc)
    e, n = bench.bcc_engine(nc, solver, 0.0, 0.5, 1, 0, 0)
    info = e.layout_info()
    assert info['runs'] == 4 and info['tile_slots'] == 1024, info
    assert np.abs(e.get_moments()[0] - S['emom']).max() <= 1e-15
    st = orc.SdState(S, solver, 1e-16, 0.5)
    e.sd_steps(10)
    for _ in range(10):
        st.step()
    err = float(np.abs(e.get_moments()[0] - st.emom).max())
    assert err <= 1e-12, (solver, err)


def test_headline_128_cubed_against_oracle_and_direct_kernel(monkeypatch):
    """The benchmark configuration itself: bcc Fe 128^3 = 4 194 304 spins, 4096 tiles, 28 waves.  (1) T = 0, midpoint, 5 steps
    against the oracle to 1e-12; (2) 300 K, both solvers: the run-compressed kernel against the direct-gather kernel
    (ASD_STAGED=0: no tiles, no run table, plain 256-bit gathers) with the same counter-based noise -- 1e-12."""
    import bench
    nc = (128, 128, 128)
    S = bench.oracle_bcc(nc)
    e, n = bench.bcc_engine(nc, 1, 0.0, 0.5, 1, 0, 0)
    assert e.layout_info()['runs'] == 4 and e.layout_info()['tile_slots'] == 1024
    st = orc.SdState(S, 1, 1e-16, 0.5)
    e.sd_steps(5)
    for _ in range(5):
        st.step()
    err = float(np.abs(e.get_moments()[0] - st.emom).max())
    assert err <= 1e-12, err
    del S, st
    e.close()
    for solver in (1, 5):
        monkeypatch.delenv('ASD_STAGED', raising=False)
        er, _ = bench.bcc_engine(nc, solver, 300.0, 0.5, 1, 0, 0)
        monkeypatch.setenv('ASD_STAGED', '0')
        ed, _ = bench.bcc_engine(nc, solver, 300.0, 0.5, 1, 0, 0)
        assert er.layout_info()['runs'] == 4 and ed.layout_info()['staged'] == 0
        a0 = er.get_moments()[0]
        er.sd_steps(8)
        ed.sd_steps(8)
        a, b = er.get_moments()[0], ed.get_moments()[0]
        assert np.abs(a - b).max() <= 1e-12, (solver, np.abs(a - b).max())
        assert np.abs(a - a0).max() > 1e-6                      # the spins did move
        er.close(); ed.close()


def test_moment_planes_follow_every_writer_of_the_spins(monkeypatch):
    """MM instantiation of the run kernel (gather list staged from the moment planes emomM[M][3][Npad], asd_runs.cuh): the planes
    are rebuilt whenever something other than an MM stage launch wrote the spins.  Same calls on an engine built with ASD_MM=0
    (gather list staged from the spins through registers): bit-identical states after steps, Monte Carlo sweeps, a new
    asd_set_moments, a fixed-moment interlude (direct kernel) and the stage-timing entry point."""
    fx, _, _ = load_golden('bccfe_cuda')
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], ncell=(64, 4, 8), mensemble=2, do_reduced='Y')
    S = orc.build_system(*args)
    inp = args[0]
    rng = np.random.default_rng(11)
    e0 = rng.normal(size=(3, S['Natom'], 2)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    e1 = rng.normal(size=(3, S['Natom'], 2)); e1 /= np.sqrt((e1 ** 2).sum(axis=0))
    e1 = np.asfortranarray(e1)
    red = np.arange(1, S['Natom'] + 1)[np.arange(S['Natom']) % 5 != 2]
    for solver in (1, 5):
        monkeypatch.delenv('ASD_MM', raising=False)
        a = _bcc_engine(S, inp, args, solver, 300.0)
        monkeypatch.setenv('ASD_MM', '0')
        b = _bcc_engine(S, inp, args, solver, 300.0)
        monkeypatch.delenv('ASD_MM', raising=False)
        assert a.layout_info()['runs'] == 4 and b.layout_info()['runs'] == 4
        assert a.layout_info()['planes'] == 1 and b.layout_info()['planes'] == 0
        for e in (a, b):
            e.sd_steps(7, first_step=1)
        assert np.array_equal(a.get_moments()[0], b.get_moments()[0]), (solver, 'steps')
        for e in (a, b):
            e.mc_sweeps('M', 3, 300.0)
            e.sd_steps(5, first_step=8)
        assert np.array_equal(a.get_moments()[0], b.get_moments()[0]), (solver, 'after Monte Carlo sweeps')
        for e in (a, b):
            e.set_moments(e1, S['mmom'])
            e.sd_steps(5, first_step=13)
        assert np.array_equal(a.get_moments()[0], b.get_moments()[0]), (solver, 'after asd_set_moments')
        for e in (a, b):
            e.set_evolving_atoms(red)
            e.sd_steps(3, first_step=18)
            e.set_evolving_atoms(None)
            e.sd_steps(4, first_step=21)
        assert np.array_equal(a.get_moments()[0], b.get_moments()[0]), (solver, 'after a fixed-moment interlude')
        for e in (a, b):
            e.time_sd_steps(0, first_step=25, stages=True)
            e.sd_steps(2, first_step=26)
        assert np.array_equal(a.get_moments()[0], b.get_moments()[0]), (solver, 'after the stage-timing entry')
        a.close(); b.close()


def test_run_kernel_with_anisotropy_and_field_against_oracle(monkeypatch):
    """Heisenberg + single-ion anisotropy (uniaxial on one sublattice, uniaxial + cubic `type 7` on the other) + a uniform field on a
    bcc lattice whose layout takes the moment-plane instantiation of the run kernel with the LEAN integrator loop (asd_runs.cuh):
    both solvers against the oracle at T = 0 to 1e-12 (hamiltonianactions.f90:225-239,842-919), and at 300 K bit for bit against
    the general instantiation (ASD_MM=0) with the same noise stream."""
    import json
    import os
    from util import GOLDEN, lattice_engine
    with open(os.path.join(GOLDEN, 'bccfe_cuda.json')) as fh:
        fx = json.load(fh)
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], ncell=(64, 8, 8), mensemble=2, do_reduced='Y', hfield=(0.3, -0.2, 1.1))
    args[9] = (np.array([1, 7], dtype=np.int32),
               np.asfortranarray([[0.12, 0.03, 0.0, 0.6, 0.8, 0.0], [-0.08, 0.05, 1.0, 0.0, 0.0, 0.35]]))
    S = orc.build_system(*args)
    inp = args[0]
    rng = np.random.default_rng(17)
    e0 = rng.normal(size=(3, S['Natom'], 2)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    assert S['aniso'] is not None and np.abs(S['external_field']).max() > 0
    for solver in (1, 5):
        monkeypatch.delenv('ASD_MM', raising=False)
        e = lattice_engine(args, S, solver, inp['timestep'], 0.3)
        info = e.layout_info()
        assert info['runs'] == 4 and info['tile_slots'] == 1024 and info['planes'] == 1, info
        beff, _ = e.effective_field()
        rb, _ = orc.effective_field(S)
        assert np.abs(beff - rb).max() <= 1e-12 * np.abs(rb).max()
        st = orc.SdState(S, solver, inp['timestep'], 0.3)
        e.sd_steps(30)
        for _ in range(30):
            st.step()
        assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12, solver
        a = lattice_engine(args, S, solver, inp['timestep'], 0.3, temp=300.0)
        monkeypatch.setenv('ASD_MM', '0')
        b = lattice_engine(args, S, solver, inp['timestep'], 0.3, temp=300.0)
        assert a.layout_info()['planes'] == 1 and b.layout_info()['planes'] == 0 and b.layout_info()['runs'] == 4
        a.sd_steps(20); b.sd_steps(20)
        assert np.array_equal(a.get_moments()[0], b.get_moments()[0]), solver
        for x in (e, a, b):
            x.close()


def test_fcc_four_sublattices_take_the_run_kernel(monkeypatch):
    """fcc in its conventional cell (four basis atoms, two shells, z = 12 + 6): the super-brick holds 2048 slots and the run kernel
    works on 1024-slot tiles (half a super-brick).  Both solvers against the oracle at T = 0 to 1e-12; at 300 K the same noise
    stream as the staged one-atom-per-thread kernel (ASD_RUNS=0), which sums the neighbours in list order: 1e-9 after 20 steps."""
    import copy
    import json
    import os
    from util import GOLDEN, lattice_engine
    with open(os.path.join(GOLDEN, 'bccfe_cuda.json')) as fh:
        fx = copy.deepcopy(json.load(fh))
    fx['posfile'] = [['1', '1', '0.0', '0.0', '0.0'], ['2', '1', '0.5', '0.5', '0.0'], ['3', '1', '0.5', '0.0', '0.5'], ['4', '1', '0.0', '0.5', '0.5']]
    fx['momfile'] = [[str(i), '1', '1.7', '0.3', '0.4', '0.85', '2.0'] for i in (1, 2, 3, 4)]
    fx['jfile'] = [['1', '1', '0.5', '0.5', '0.0', '1.2'], ['1', '1', '1.0', '0.0', '0.0', '-0.15']]
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], ncell=(64, 8, 8), mensemble=2, do_reduced='Y', hfield=(0.0, 0.0, 0.5))
    S = orc.build_system(*args)
    inp = args[0]
    assert S['Natom'] == 4 * 64 * 8 * 8 and S['exchange']['z'] == 18 and S['nHam'] == 4
    rng = np.random.default_rng(23)
    e0 = rng.normal(size=(3, S['Natom'], 2)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    for solver in (1, 5):
        monkeypatch.delenv('ASD_RUNS', raising=False)
        e = lattice_engine(args, S, solver, inp['timestep'], 0.3)
        info = e.layout_info()
        assert info['runs'] == 4 and info['tile_slots'] == 1024, info
        beff, _ = e.effective_field()
        rb, _ = orc.effective_field(S)
        assert np.abs(beff - rb).max() <= 1e-12 * np.abs(rb).max()
        st = orc.SdState(S, solver, inp['timestep'], 0.3)
        e.sd_steps(30)
        for _ in range(30):
            st.step()
        assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12, solver
        a = lattice_engine(args, S, solver, inp['timestep'], 0.3, temp=300.0)
        monkeypatch.setenv('ASD_RUNS', '0')
        b = lattice_engine(args, S, solver, inp['timestep'], 0.3, temp=300.0)
        assert b.layout_info()['runs'] == 0 and b.layout_info()['staged'] == 1
        a.sd_steps(20); b.sd_steps(20)
        assert np.abs(a.get_moments()[0] - b.get_moments()[0]).max() <= 1e-9, solver
        for x in (e, a, b):
            x.close()


@pytest.mark.parametrize('what', ['per_site_damping', 'site_field', 'torque', 'mompar'])
def test_general_instantiation_on_moment_planes(what, monkeypatch):
    """A moment-plane layout whose run takes the GENERAL integrator loop (llg_runs_kernel<.., LEAN 0, MM>): per-site damping, a
    site-dependent external field, a spin-transfer-torque field, or mompar 1 switch the LEAN flavour off at launch time.  Both
    solvers against the oracle at T = 0 to 1e-12."""
    from util import fixture_args, lattice_engine
    args = fixture_args('bccfe_cuda', mens=2, ncell=(64, 8, 8), do_reduced='Y')
    S = orc.build_system(*args)
    inp = args[0]
    rng = np.random.default_rng(29)
    N = S['Natom']
    e0 = rng.normal(size=(3, N, 2)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    damping = 0.3
    mompar = 0
    bt = None
    if what == 'per_site_damping':
        damping = 0.05 + 0.4 * rng.random(N)
    if what == 'site_field':
        S['external_field'] = np.asfortranarray(rng.normal(size=(3, N, 2)) * 2.0)
    if what == 'torque':
        bt = np.asfortranarray(rng.normal(size=(3, N, 2)) * 0.5)
    if what == 'mompar':
        mompar = 1
    for solver in (1, 5):
        from uppasd_b200 import host, lattice
        c = orc.consts(S)
        e = host.Engine()
        e.set_constants(c['gama'], c['k_bolt'], c['mub'], c['mry'])
        e.set_system(N, 2, S['nHam'], S['aHam'])
        nn, red, xc, nntype = args[6](S)
        ns, ca, cs, sh = lattice.stencil(inp['cell'], S['bas'], S['atype_inp'], nn, red, inp['sym'], nntype, ncell=inp['ncell'])
        cp = lattice.couplings(ns, ca, sh, S['atype_inp'], xc, S['ammom_inp'], c['mry'], c['mub'], 1)
        e.build_lattice_table(0, S['NA'], inp['ncell'], inp['bc'], ns, ca, cs, cp)
        e.set_external_field(S['external_field'])
        if bt is not None:
            e.set_torque(bt)
        e.set_llg(solver, inp['timestep'], landeg=S['Landeg'], lambda1=damping, temp=0.0, mompar=mompar)
        e.set_moments(S['emom'], S['mmom'], S['mmom0'])
        e.commit()
        assert e.layout_info()['planes'] == 1 and e.layout_info()['runs'] == 4
        st = orc.SdState(S, solver, inp['timestep'], damping, mompar=mompar, btorque=bt)
        e.sd_steps(25)
        for _ in range(25):
            st.step()
        emom, emomM, mmom = e.get_moments()
        assert np.abs(emom - st.emom).max() <= 1e-12, (what, solver, np.abs(emom - st.emom).max())
        assert np.abs(mmom - st.mmom).max() <= 1e-12 * np.abs(st.mmom).max(), (what, solver)
        e.close()
