"""The reference's own regression cases, run the way a user runs them: a copy of the reference run directory
(materialised from tests/golden/*.json) -> `uppasd_b200.driver.Simulation(...).run()` -> the reference's measurement
files, read back row by row like tests/bergtest.py does and compared with the values its YAML files pin
(tests/regulartests.yaml, tests/regressionResaro.yaml, tests/cudatests.yaml; tolerances = bergtest's)."""
import json
import os

import numpy as np
import pytest

from oracle import inputs, orc
from test_asdio_host import materialise
from uppasd_b200 import asdio
from util import GOLDEN

pytestmark = pytest.mark.gpu


def _row(path, it):
    for r in asdio.read_out(path):
        if int(r[0]) == it:
            return r
    raise AssertionError('no row %d in %s' % (it, path))


def _sloppy(a, b):
    return abs(a - b) <= 2e-2 and abs(a - b) <= 2e-2 * max(abs(a), abs(b), 1e-300)


def test_kagome_run_directory(tmp_path):
    from uppasd_b200 import driver
    fx, path = materialise('kagome', tmp_path)
    sim = driver.Simulation(path).run()
    d = str(tmp_path)
    exp = fx['expected']
    r = _row(os.path.join(d, 'averages.kagome_T.out'), 13000)
    for a, b in zip(r[1:5], exp['averages']['13000']):
        assert abs(a - b) <= 1e-8                                       # regulartests.yaml:132-143 ('similar')
    t = _row(os.path.join(d, 'trajectory.kagome_T.002.1.out'), 2400)
    assert int(t[1]) == 2
    for a, b in zip(t[2:6], exp['trajectory']['2400']):
        assert abs(a - b) <= 1e-8                                       # regulartests.yaml:145-156
    # restart file written at the end (uppasd.f90:344-350), readable by initmag 4
    rstep, emom, mmom = asdio.read_restart(os.path.join(d, 'restart.kagome_T.out'), sim.natom, 1)
    assert rstep == 15001 and np.abs(emom - sim.moments()[0]).max() < 1e-8
    assert os.path.exists(os.path.join(d, 'totenergy.kagome_T.out')) and os.path.exists(os.path.join(d, 'coord.kagome_T.out'))


def test_megatest_run_directory(tmp_path):
    from uppasd_b200 import driver
    fx, path = materialise('megatest', tmp_path)
    sim = driver.Simulation(path).run()
    d = str(tmp_path)
    exp = fx['expected']
    r = _row(os.path.join(d, 'averages.megaTest.out'), 11000)
    for a, b in zip(r[1:5], exp['averages']['11000']):
        assert abs(a - b) <= 1e-8                                       # regressionResaro.yaml:18-27
    emom, _, _ = sim.moments()
    for a, b in zip(emom[:, exp['moment']['atom'] - 1, 0], exp['moment']['11000']):
        assert abs(a - b) <= 1e-8                                       # regressionResaro.yaml:64-73
    c = asdio.read_out(os.path.join(d, 'coord.megaTest.out'))[exp['coord']['atom'] - 1]
    assert c[1:4] == exp['coord']['row'] and int(c[4]) == exp['coord']['type'] and int(c[5]) == exp['coord']['numb']
    # totenergy row 10900 (regressionResaro.yaml:112-122): Tot, Exc, ..., Zeeman
    e = _row(os.path.join(d, 'totenergy.megaTest.out'), 10900)
    assert abs(e[1] - (-6.10005366)) <= 1e-8
    assert abs(e[9] - (-5.36620122e-05)) <= 1e-8
    assert abs(e[1] - (e[2] + e[9])) <= 3e-8                              # printed with es16.8
    # cumulant row 211 with the energy running means (regressionResaro.yaml:112-122) and the per-site projected averages
    # (:168-179, bergtest `almost` = 5e-4 in the reference's YAML)
    cu = [x for x in asdio.read_out(os.path.join(d, 'cumulants.megaTest.out')) if int(x[0]) == 211][0]
    for a, b in zip(cu[1:9], exp['cumulants']['211']):
        assert abs(a - b) <= 5e-9 * max(1.0, abs(b)), (cu, exp['cumulants'])
    pr = [x for x in asdio.read_out(os.path.join(d, 'projavgs.megaTest.out')) if int(x[0]) == 11000 and int(x[1]) == 2][0]
    for a, b in zip(pr[2:7], exp['projavgs']['11000']['2']):
        assert abs(a - b) <= 5e-4, (pr, exp['projavgs'])
    assert abs(pr[4] - 2.5 * exp['moment']['11000'][0]) <= 1e-8
    # projcumulants row 211 of type 2 (regressionResaro.yaml:181-190)
    pc = [x for x in asdio.read_out(os.path.join(d, 'projcumulants.megaTest.out')) if int(x[0]) == 211 and int(x[1]) == 2][0]
    for a, b in zip(pc[1:7], [2, 2.5, 6.25, 39.0625, 0.666666667, -7.08541485e-37]):
        assert abs(a - b) <= 1e-8, pc


@pytest.mark.parametrize('name,simid,sdealgh', [('feco', 'FeCo__B2', 1), ('feco_cuda', 'FeCo__B2', 5), ('bccfe_cuda', 'bcc_Fe_T', 5)])
def test_cumulant_goldens(name, simid, sdealgh, tmp_path):
    """tests/FeCo (regulartests.yaml:219-242) and the reference's CUDA cases (cudatests.yaml:24-70; its CUDA path always
    integrates with Depondt, so SDEalgh 5 is requested explicitly here)"""
    from uppasd_b200 import driver
    fx, path = materialise(name, tmp_path)
    inp = asdio.read_inpsd(path)
    inp['sdealgh'] = sdealgh
    driver.Simulation(inp, directory=str(tmp_path)).run()
    exp = fx['expected']
    (it, want), = exp['averages_M'].items()
    assert _sloppy(_row(os.path.join(str(tmp_path), 'averages.%s.out' % simid), int(it))[4], want)
    (it, want), = exp['cumulants'].items()
    got = _row(os.path.join(str(tmp_path), 'cumulants.%s.out' % simid), int(it))
    for a, b in zip(got[1:5], want):
        assert _sloppy(a, b), (got, want)


@pytest.mark.parametrize('mode', ['S', 'M', 'H'])
def test_thermal_bccfe_modes_are_statistically_right(mode, tmp_path):
    """tests/bccFe (6^3, 500 K, initial phase + measurement phase in modes S / M / H).  Its goldens pin the reference's
    own random stream (tseed 5, one thread); with a different generator the same observables must agree within a
    few per cent: <M> ~ 1.80 mu_B, U_Binder ~ 0.666 (regulartests.yaml:244-347)."""
    from uppasd_b200 import driver
    fx, path = materialise('bccfe', tmp_path, {'MODE': mode})
    inp = asdio.read_inpsd(path)
    inp['mcnstep'] = 3000 if mode != 'S' else 0
    inp['mensemble'] = 4
    driver.Simulation(inp, directory=str(tmp_path)).run()
    rows = asdio.read_out(os.path.join(str(tmp_path), 'cumulants.bcc_Fe_T.out'))
    last = rows[-1]
    ref = fx['expected']['S_cumulants_41']
    assert abs(last[1] - ref[0]) < 0.06, (mode, last, ref)               # <M>
    assert abs(last[4] - ref[3]) < 5e-3                                  # Binder cumulant deep in the ordered phase
    av = asdio.read_out(os.path.join(str(tmp_path), 'averages.bcc_Fe_T.out'))
    assert abs(av[-1][4] - 1.80) < 0.08


def test_restart_and_relax_api(tmp_path):
    """initmag 4 continues a run from restart.<simid>.out (restart.f90:320-381); relax() is pyasd's relax_"""
    from uppasd_b200 import driver
    fx, path = materialise('kagome', tmp_path)
    inp = asdio.read_inpsd(path)
    inp['nstep'] = 400
    a = driver.Simulation(dict(inp), directory=str(tmp_path)).run()
    inp2 = dict(inp, initmag=4, restartfile=os.path.join(str(tmp_path), 'restart.kagome_T.out'), nstep=200)
    b = driver.Simulation(inp2, directory=str(tmp_path))
    assert b.rstep == 401
    assert np.abs(b.moments()[0] - a.moments()[0]).max() < 1e-8           # es16.8 round trip
    m = b.relax('S', 50, 0.0, 1e-16, 0.5)
    # unit norm to the precision of the es16.8 restart file (the schemes conserve whatever norm they are given)
    assert m.shape == (3, b.natom, 1) and np.abs(np.sqrt((m ** 2).sum(axis=0)) - 1).max() < 1e-8
    e0 = b.energy()
    b.relax('H', 20, 1.0)
    b.relax('S', 200, 0.0, 1e-16, 0.5)
    assert b.energy() <= e0 + 1e-6


def test_cluster_run_directory(tmp_path):
    """tests/Cluster through the driver (regulartests.yaml:182-205, tol 1e-8): random start from the reference's own
    generator (uppasd_b200/refrng.py), two Depondt initial phases, BQ + type-7 anisotropy, energy columns."""
    from uppasd_b200 import driver
    fx, path = materialise('cluster', tmp_path)
    driver.Simulation(path).run()
    d, exp = str(tmp_path), fx['expected']
    r = _row(os.path.join(d, 'averages.ClusterT.out'), 25000)
    for a, b in zip(r[1:5], exp['averages']['25000']):
        assert abs(a - b) <= 1e-8, (r, exp)
    e = _row(os.path.join(d, 'totenergy.ClusterT.out'), 15000)
    want = exp['totenergy']['15000']
    for col, key in ((1, 'tot'), (2, 'exc'), (3, 'ani'), (7, 'bq')):
        assert abs(e[col] - want[key]) <= 1e-8, (key, e, want)


def test_heischainaf_run_directory_projected_averages(tmp_path):
    """tests/HeisChainAF through the driver (regulartests.yaml:54-103): averages, total energy, and projavgs.*.out from the
    device's sublattice sums (asd_measure_sublattice)."""
    from uppasd_b200 import driver
    fx, path = materialise('heischainaf', tmp_path)
    inp = asdio.read_inpsd(path)
    inp['nstep'] = 5100
    driver.Simulation(inp, directory=str(tmp_path)).run()
    d, exp = str(tmp_path), fx['expected']
    r = _row(os.path.join(d, 'averages.AF_WireT.out'), 1000)
    for a, b in zip(r[1:5], exp['averages']['1000']):
        assert abs(a - b) <= 1e-8
    assert abs(_row(os.path.join(d, 'totenergy.AF_WireT.out'), 1500)[1] - exp['totenergy']['1500']['tot']) <= 1e-8
    rows = [x for x in asdio.read_out(os.path.join(d, 'projavgs.AF_WireT.out')) if int(x[0]) == 5000]
    assert len(rows) == 2
    for x in rows:
        want = exp['projavgs']['5000'][str(int(x[1]))]
        for a, b in zip((x[2], x[4], x[5], x[6]), want):
            assert abs(a - b) <= 1e-8, (x, want)


@pytest.mark.parametrize('name,simid', [('heisstripe', 'HeisStri'), ('scsurf', 'SCsurf_T')])
def test_more_run_directories(name, simid, tmp_path):
    """tests/HeisStripe (uniaxial anisotropy, open edges) and tests/SCsurf (aunits Y, DM, anisotropy; its `skyno Y` is not
    on this path and only warns) through the driver, reference tolerances (regulartests.yaml:105-129, 158-170)."""
    import warnings
    from uppasd_b200 import driver
    fx, path = materialise(name, tmp_path)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        driver.Simulation(path).run()
    d, exp = str(tmp_path), fx['expected']
    for it, want in exp['averages'].items():
        r = _row(os.path.join(d, 'averages.%s.out' % simid), int(it))
        for a, b in zip(r[1:5], want):
            assert abs(a - b) <= 1e-8, (name, r, want)
    for it, want in exp.get('totenergy', {}).items():
        e = _row(os.path.join(d, 'totenergy.%s.out' % simid), int(it))
        for col, key in ((1, 'tot'), (2, 'exc'), (3, 'ani')):
            assert abs(e[col] - want[key]) <= 1e-8, (name, key, e, want)


def test_kagome_tensor_run_directory(tmp_path):
    """tests/kagome_cuda through the driver (cudatests.yaml:1-23, tol 1e-8): jfile.tensor read by the product, the j_tens
    table built on the device (builder kind 3), random start from the reference's generator, Depondt (what the reference's
    CUDA path runs): averages @1300 and the cumulant row 171 from the written files."""
    from uppasd_b200 import driver
    fx, path = materialise('kagome_cuda', tmp_path)
    inp = asdio.read_inpsd(path)
    inp['sdealgh'] = 5
    driver.Simulation(inp, directory=str(tmp_path)).run()
    d, exp = str(tmp_path), fx['expected']
    r = _row(os.path.join(d, 'averages.kagome_T.out'), 1300)
    for a, b in zip(r[1:5], exp['averages']['1300']):
        assert abs(a - b) <= 1e-8, (r, exp)
    cu = [x for x in asdio.read_out(os.path.join(d, 'cumulants.kagome_T.out')) if int(x[0]) == 171][0]
    for a, b in zip(cu[1:5], exp['cumulants']['171']):
        assert abs(a - b) <= 1e-8, (cu, exp)


@pytest.mark.parametrize('kind,alg', [(1, 1), (4, 5)])
def test_field_pulse_through_the_driver(kind, alg, tmp_path):
    """do_bpulse in a run directory: the driver reads the bpulsefile, hands the schedule of the measurement phase to the engine once
    (asd_set_time_field) and the stage kernels add the pulse of their step; final state against the oracle's sd_mphase loop with the
    pulse evaluated step by step (sd_driver.f90:389-393, 703-722, 770-779), T = 0, to 1e-12"""
    from test_asdio_host import BPULSE_FILES
    from uppasd_b200 import driver, fields
    fx = json.load(open(os.path.join(GOLDEN, 'kagome.json')))
    d = tmp_path / 'run'
    d.mkdir()
    for k, v in fx['raw'].items():
        (d / k).write_text(v)
    drop = ('nstep', 'sdealgh', 'do_avrg', 'temp', 'initmag', 'mensemble')
    lines = [l for l in fx['raw']['inpsd.dat'].splitlines() if not (l.split() and l.split()[0].lower() in drop)]
    lines += ['nstep 120', 'sdealgh %d' % alg, 'do_avrg N', 'temp 0.0', 'initmag 3', 'mensemble 2', 'do_bpulse %d' % kind, 'bpulsefile ./bpulsefile']
    (d / 'inpsd.dat').write_text('\n'.join(lines) + '\n')
    (d / 'bpulsefile').write_text(BPULSE_FILES[kind])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sim = driver.Simulation(str(d / 'inpsd.dat'))
        sim.run()
    inp = sim.inp
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], mensemble=2)
    S = orc.build_system(*args)
    P = fields.read_bpulse(str(d / 'bpulsefile'), kind)
    B = orc.bpulse_setup(kind, P['b0'], P['step'], P['par'][:6])
    st, seen = orc.sd_run_bpulse(S, alg, inp['timestep'], inp['damping'], B, 0, 120)
    assert np.abs(seen).max() > 1.0                                   # the pulse is there
    assert np.abs(sim.bpulse - seen).max() <= 1e-13 * np.abs(seen).max()
    got = sim.engine.get_moments()[0]
    assert np.abs(got - st.emom).max() <= 1e-12
    assert np.abs(got - S['emom']).max() > 1e-3
