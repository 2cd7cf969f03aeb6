"""CPU checks of the product's readers / writers / estimators for the reference's on-disk formats
against the test-only restatement in oracle/ and against the reference's Fortran edit descriptors."""
import json
import os

import numpy as np
import pytest

from oracle import inputs as oinputs
from oracle import orc
from uppasd_b200 import asdio, observables
from util import GOLDEN

NAMES = ['kagome', 'megatest', 'feco', 'bccfe_cuda', 'bccfe']


def materialise(name, tmp_path, subst=None):
    fx = json.load(open(os.path.join(GOLDEN, name + '.json')))
    for rel, text in fx['raw'].items():
        for a, b in (subst or {}).items():
            text = text.replace(a, b)
        p = os.path.join(str(tmp_path), rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, 'w') as fh:
            fh.write(text)
    return fx, os.path.join(str(tmp_path), 'inpsd.dat')


@pytest.mark.parametrize('name', NAMES)
def test_inpsd_reader_agrees_with_the_restated_parser(name, tmp_path):
    fx, path = materialise(name, tmp_path, {'MODE': 'S'})
    got = asdio.read_inpsd(path)
    ref = oinputs.read_inpsd(path)
    for k in ('simid', 'ncell', 'bc', 'sym', 'posfiletype', 'maptype', 'mensemble', 'tseed', 'sdealgh', 'initmag', 'mode', 'temp',
              'nstep', 'damping', 'timestep', 'hfield', 'do_reduced', 'mompar', 'avrg_step', 'cumu_step', 'cumu_buff', 'do_cumu',
              'plotenergy', 'gpu_mode', 'ip_mode'):
        assert got[k] == ref[k] or (isinstance(ref[k], float) and abs(got[k] - ref[k]) == 0.0), (k, got[k], ref[k])
    assert np.array_equal(got['cell'], ref['cell'])


@pytest.mark.parametrize('name', NAMES)
def test_structure_readers_agree(name, tmp_path):
    fx, path = materialise(name, tmp_path, {'MODE': 'S'})
    inp = asdio.read_inpsd(path)
    ref = oinputs.read_inpsd(path)
    bas, atype = asdio.read_posfile(inp['posfile'], inp['cell'], inp['posfiletype'])
    rbas, ratype = oinputs.read_positions(ref['files']['posfile'], ref['cell'], ref['posfiletype'])
    assert np.array_equal(bas, rbas) and np.array_equal(atype, ratype)
    na = bas.shape[1]
    a, b, c = asdio.read_momfile(inp['momfile'], na)
    ra, rb, rc = oinputs.read_moments(ref['files']['momfile'], na)
    assert np.array_equal(a, ra) and np.array_equal(b, rb) and np.array_equal(c, rc)
    for key, okey, ncomp in (('exchange', 'exchange', 1), ('dm', 'dm', 3)):
        if not inp.get(key):
            continue
        nn, red, xc, nnt = asdio.read_pairfile(inp[key], atype, bas, inp['cell'], inp['maptype'], inp['posfiletype'], ncomp)
        rnn, rred, rxc, rnnt = oinputs.read_pair_file(ref['files'][okey], int(atype.max()), atype, bas, ref['cell'], ref['maptype'],
                                                      ref['posfiletype'], ncomp, True)
        assert np.array_equal(nn, rnn) and np.array_equal(red, rred) and np.array_equal(xc, rxc) and np.array_equal(nnt, rnnt)


def test_block_keywords_of_the_thermal_fixture(tmp_path):
    fx, path = materialise('bccfe', tmp_path, {'MODE': 'M'})
    inp = asdio.read_inpsd(path)
    assert inp['mode'] == 'M' and inp['ip_mode'] == 'M'
    assert inp['ip_nphase'] == [(2000, 500.0, 1.0e-16, 0.5)]
    assert inp['ip_mcanneal'] == [(2000, 500.0)]
    fx, path = materialise('kagome', tmp_path)
    assert asdio.read_inpsd(path)['trajectories'] == [(2, 100, 1)]


def test_fortran_edit_descriptors():
    assert asdio.es16_8(1.72521729) == '  1.72521729E+00'
    assert asdio.es16_8(-0.0656709236) == ' -6.56709236E-02'
    assert asdio.es16_8(0.0) == '  0.00000000E+00'
    assert asdio.es16_8(1.0e-16) == '  1.00000000E-16'
    assert asdio.es16_8(-2.5e-120) == ' -2.50000000-120'      # gfortran drops the E for three-digit exponents
    assert len(asdio.es16_8(-1.234e150)) == 16


def test_output_files_round_trip(tmp_path):
    out = asdio.OutputFiles(str(tmp_path), 'abc')
    out.averages([(0, 1.0, 2.0, 3.0, 3.7416573867739413, 0.0)])
    out.averages([(100, 0.5, -0.25, 0.125, 0.57282196186948, 1e-3)])
    rows = asdio.read_out(os.path.join(str(tmp_path), 'averages.abc.out'))
    assert rows[1][0] == 100 and abs(rows[1][2] + 0.25) < 1e-15
    lines = open(os.path.join(str(tmp_path), 'averages.abc.out')).read().splitlines()
    assert lines[0].split() == ['#Iter', '<M>_x', '<M>_y', '<M>_z', '<M>', 'M_{stdv}'] and len(lines[1]) == 8 + 5 * 16
    emom = np.zeros((3, 2, 1), order='F'); emom[2] = 1.0
    out.restart(7, 'S', emom, np.full((2, 1), 2.23, order='F'))
    rstep, e2, m2 = asdio.read_restart(os.path.join(str(tmp_path), 'restart.abc.out'), 2, 1)
    assert rstep == 7 and np.array_equal(e2, emom) and np.allclose(m2, 2.23)
    out.trajectory(2, 1, [(0, 0.1, 0.2, 0.3, 2.0)])
    assert os.path.exists(os.path.join(str(tmp_path), 'trajectory.abc.002.1.out'))


def test_cumulant_estimator_matches_the_restatement():
    rng = np.random.default_rng(3)
    n, m = 432, 3
    mine = observables.Cumulants(n, m, 300.0, orc.CONST['k_bolt'], orc.CONST['mub'], orc.CONST['mry'], buff=1)
    ref = orc.Cumulants(n)
    last = None
    for s in range(40):
        msum = rng.normal(size=(3, m)) * 50 + np.array([[0.0], [0.0], [700.0]])
        row = mine.sample(msum)
        ref.sample(msum)
        last = row
    got = np.array(last[1:5])
    want = np.array([ref.m1, ref.m2, ref.m4, ref.binder])
    assert np.allclose(got, want, rtol=1e-13, atol=0)
    assert last[0] == 40 and mine.chi > 0.0


def test_averages_buffer():
    av = observables.Averages(10, buff=2)
    assert av.sample(0, np.array([[1.0, 3.0], [0.0, 0.0], [0.0, 4.0]])) is None
    rows = av.sample(100, np.array([[2.0, 2.0], [0.0, 0.0], [0.0, 0.0]]))
    assert len(rows) == 2 and rows[0][0] == 0 and abs(rows[0][4] - 0.5 * (0.1 + 0.5)) < 1e-15 and rows[1][5] == 0.0
