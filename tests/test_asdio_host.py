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

NAMES = ['kagome', 'megatest', 'feco', 'bccfe_cuda', 'bccfe', 'cluster', 'heisstripe', 'heischainaf', 'scsurf']


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


def test_reference_uniform_generator_and_random_start():
    """uppasd_b200/refrng.py (product, Python integers) against the reference's known answers and
    against the test-only C++ restatement: Initmag 1 starts are identical bit for bit."""
    from uppasd_b200 import refrng
    from util import load_golden
    g = refrng.ReferenceUniform(5)
    assert [g.word() for _ in range(3)] == [953453411, 236996814, 2970113047]
    for name in ('solvers', 'cluster', 'heisstripe'):
        fx, inp, S = load_golden(name)
        orc.initmag1(S, inp['tseed'])
        e = refrng.random_start(S['NA'], S['ncell'], inp['tseed'])
        assert np.array_equal(e, S['emom'][:, :, 0]), name


def test_triangulation_and_projected_estimators():
    from uppasd_b200 import lattice
    from util import load_golden
    for dims in ((3, 2, 1, 1), (5, 4, 2, 3), (16, 16, 1, 1)):
        assert np.array_equal(lattice.triangulation(*dims), orc.delaunay_tri_tri(*dims))
    # projavgs rows from the oracle's state of tests/HeisChainAF at iteration 5000 (regulartests.yaml:69-92)
    fx, inp, S = load_golden('heischainaf')
    orc.initmag1(S, inp['tseed'])
    st = orc.SdState(S, inp['sdealgh'], inp['timestep'], inp['damping'])
    for _ in range(5000):
        st.step()
    na = S['NA']
    msum_na = np.stack([st.emomM[:, c::na, :].sum(axis=1) for c in range(na)], axis=1)       # (3, NA, M)
    rows = observables.projected_rows(5000, msum_na, S['Natom'] // na, S['atype_inp'], 'Y')
    for r in rows:
        want = fx['expected']['projavgs']['5000'][str(r[1])]
        for a, b in zip((r[2], r[4], r[5], r[6]), want):
            assert abs(a - b) <= 1e-8, (r, want)
    # running mean / variance of the skyrmion number (prn_topology.f90:321-327)
    sk = observables.SkyrmionNumber(1)
    xs = [1.0, 0.5, 2.0, -1.0]
    for i, x in enumerate(xs):
        row = sk.sample(i, [x, x])
    assert abs(row[2] - np.mean(xs)) < 1e-15 and abs(row[3] - np.var(xs)) < 1e-15


def test_skyrmion_number_of_a_neel_skyrmion_is_an_integer():
    n = 32
    a1, a2 = np.array([1.0, 0.0]), np.array([-0.5, 0.866025403784])
    ix, iy = np.meshgrid(np.arange(n), np.arange(n), indexing='xy')
    pos = ix.ravel()[:, None] * a1 + iy.ravel()[:, None] * a2
    d = pos - pos.mean(axis=0)
    r, phi = np.hypot(d[:, 0], d[:, 1]), np.arctan2(d[:, 1], d[:, 0])
    th = np.pi * np.exp(-r / 4.0)
    e = np.stack([np.sin(th) * np.cos(phi), np.sin(th) * np.sin(phi), np.cos(th)])[:, :, None]
    q, per = orc.pontryagin_tri(e, orc.delaunay_tri_tri(n, n, 1, 1))
    assert abs(q + 1.0) < 1e-12 and abs(per[0] + 1.0) < 1e-12


def test_single_cell_stencil_keeps_folded_neighbours():
    """tests/Cluster is ONE cell with 43 basis atoms and BC 0 0 0; several basis positions are negative and get folded to the
    far side of the 10 x 10 x 10 cell.  The reference keeps the neighbours found through the fold because a single cell
    has no cell hops at all (neighbourmap.f90:226-230, 270-296).  The product's stencil must do the same: its entries,
    de-duplicated in order (hamiltonianinit.f90:1055-1059), are the oracle's neighbour lists."""
    from uppasd_b200 import lattice
    from util import load_golden
    fx, inp, S = load_golden('cluster')
    args = oinputs.load_fixture(fx)
    for key, mk, sym, typed in (('exchange', args[6], inp['sym'], True), ('bq', args[8], inp['sym'], False)):
        nn, red, xc, nntype = mk(S)
        ns, ca, cs, sh = lattice.stencil(inp['cell'], S['bas'], S['atype_inp'], nn, red, sym, nntype if typed else None, ncell=inp['ncell'])
        assert not cs.any()
        for i0 in range(S['NA']):
            seen = []
            for q in range(ns[i0]):
                if ca[i0, q] not in seen:
                    seen.append(int(ca[i0, q]))
            n = S[key]['listsize'][i0]
            assert seen == list(S[key]['list'][:n, i0]), (key, i0)


def test_tensor_exchange_reader_agrees(tmp_path):
    """jfile.tensor of tests/kagome_cuda: the product's reader (row-by-row tensor, transposed into Fortran storage order)
    against the restated read_exchange_tensor_base."""
    fx, path = materialise('kagome_cuda', tmp_path)
    inp, ref = asdio.read_inpsd(path), oinputs.read_inpsd(path)
    assert inp['do_jtensor'] == 1
    bas, atype = asdio.read_posfile(inp['posfile'], inp['cell'], inp['posfiletype'])
    a = asdio.read_tensorfile(inp['exchange'], atype, bas, inp['cell'], inp['maptype'], inp['posfiletype'])
    b = oinputs.read_tensor_file(ref['files']['exchange'], int(atype.max()), atype, bas, ref['cell'], ref['maptype'], ref['posfiletype'])
    assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3] is None
    assert np.abs(a[2]).max() > 0 and not np.array_equal(a[2][1], a[2][3])      # an asymmetric tensor: the transpose matters


def test_projected_cumulants_estimator_and_file(tmp_path):
    """projcumulants.megaTest.out row 211 of type 2 (regressionResaro.yaml:181-190): the product's per-type estimator (plain
    running means, calc_and_print_cumulant_proj) fed with oracle states, written and read back in the reference's format."""
    from util import load_golden
    fx, inp, S = load_golden('megatest')
    N, na, c = S['Natom'], S['NA'], orc.consts(S)
    st = orc.SdState(S, 1, inp['timestep'], inp['damping'])
    pc = observables.ProjectedCumulants(S['atype_inp'], N // na, 1, inp['temp'], c['k_bolt'], c['mub'], inp['cumu_buff'])
    out = asdio.OutputFiles(str(tmp_path), 'megaTest')
    for mstep in range(1, 10551):
        if mstep % inp['cumu_step'] == 0:
            rows = pc.sample(np.stack([st.emomM[:, q::na, :].sum(axis=1) for q in range(na)], axis=1))
            if rows:
                out.projcumulants(rows)
        st.step()
    got = [r for r in asdio.read_out(os.path.join(str(tmp_path), 'projcumulants.megaTest.out')) if int(r[0]) == 211 and int(r[1]) == 2][0]
    for a, b in zip(got[1:7], [2, 2.5, 6.25, 39.0625, 0.666666667, -7.08541485e-37]):
        assert abs(a - b) <= 1e-8, got


def test_input_edge_cases(tmp_path):
    """Fortran-style numbers (`1.0d-16`, `.5`), logical flags written `.true.` / `T`, comment characters, unknown keywords,
    a truncated multi-row block, and empty / ragged data files."""
    d = str(tmp_path)
    with open(os.path.join(d, 'posfile'), 'w') as fh:
        fh.write('# comment line\n1 1 0.0 0.0 0.0\n\n2 1 .5 .5 0.5d0   trailing words are ignored\n')
    with open(os.path.join(d, 'momfile'), 'w') as fh:
        fh.write('1 1 2.0d0 0 0 1\n2 1 1.5 0.0 3.0 4.0\n')
    with open(os.path.join(d, 'jfile'), 'w') as fh:
        fh.write('1 1 1.0 0.0 0.0 1.0d0\n1 1 1.0 0.0 0.0 2.0\n')          # the later line overwrites the shell's value
    with open(os.path.join(d, 'inpsd.dat'), 'w') as fh:
        fh.write('simid averylongname\n% a comment\nncell 2 3 4\nBC P 0 p\ncell 1 0 0\n 0 1.0d0 0\n\n 0 0 1\n'
                 'posfile ./posfile\nmomfile ./momfile\nexchange ./jfile\nsome_unknown_keyword 1 2 3\n'
                 'timestep 1.0d-16\ndamping .5\nmap_multiple .true.\ndo_cumu y\nTemp 3.0D2\nip_nphase 2\n100 1.0 1d-16 0.1\n\n200 2.0 2d-16 0.2\n')
    inp = asdio.read_inpsd(os.path.join(d, 'inpsd.dat'))
    assert inp['simid'] == 'averylon' and inp['ncell'] == (2, 3, 4) and inp['bc'] == ('P', '0', 'P')
    assert inp['timestep'] == 1.0e-16 and inp['damping'] == 0.5 and inp['temp'] == 300.0 and inp['map_multiple'] is True
    assert inp['do_cumu'] == 'Y' and inp['ip_nphase'] == [(100, 1.0, 1e-16, 0.1), (200, 2.0, 2e-16, 0.2)]
    assert np.array_equal(inp['cell'], np.eye(3))
    bas, atype = asdio.read_posfile(inp['posfile'], inp['cell'], 'C')
    assert bas.shape == (3, 2) and np.array_equal(bas[:, 1], [0.5, 0.5, 0.5]) and list(atype) == [1, 1]
    ammom, aemom, _ = asdio.read_momfile(inp['momfile'], 2)
    assert list(ammom) == [2.0, 1.5] and np.allclose(aemom[:, 1], [0.0, 0.6, 0.8])
    nn, red, xc, nnt = asdio.read_pairfile(inp['exchange'], atype, bas, inp['cell'], 1, 'C', 1)
    assert list(nn) == [1] and xc[0, 0, 0] == 2.0
    # a block that ends early is an input error, not a silent truncation
    with open(os.path.join(d, 'bad.dat'), 'w') as fh:
        fh.write('ip_nphase 2\n100 1.0 1d-16 0.1\n')
    with pytest.raises(asdio.InputError):
        asdio.read_inpsd(os.path.join(d, 'bad.dat'))
    # empty data file: no shells
    open(os.path.join(d, 'empty'), 'w').close()
    nn, red, xc, nnt = asdio.read_pairfile(os.path.join(d, 'empty'), atype, bas, inp['cell'], 1, 'C', 1)
    assert list(nn) == [0]


def test_inpsd_keywords_outside_the_path_are_not_silent(tmp_path):
    """a keyword that switches on physics this path does not implement must not run to completion with reference-
    format files.  read_inpsd records Hamiltonian / dynamics keywords in `unserved` (driver.Simulation raises Unsupported),
    measurements it does not write in `unwritten`, unknown keywords in `ignored`; keywords at their 'off' value are fine."""
    from uppasd_b200 import asdio, driver
    p = tmp_path / 'inpsd.dat'
    p.write_text('simid x\nncell 2 2 2\ndo_dip 1\nstt A\ndo_bpulse 0\ndo_lsf N\nchir ./chirfile\ndo_sc C\ndo_ams N\nfoo_bar 3\n'
                 'do_cumu A\nskyno Y\ngradtemp 1\nmult_axis N\n')
    d = asdio.read_inpsd(str(p))
    assert [k for k, _ in d['unserved']] == ['do_dip', 'stt', 'chir', 'gradtemp']
    assert d['unwritten'] == ['do_sc', 'do_cumu A', 'skyno Y'] and d['ignored'] == ['foo_bar']
    with pytest.raises(driver.Unsupported, match='do_dip 1, stt A'):
        driver.Simulation(d, directory=str(tmp_path))
    # every run directory of the reference that this path serves stays accepted (measurement-only extras are warnings)
    import glob
    import json
    for f in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', '*.json')):
        raw = json.load(open(f)).get('raw', {})
        if 'inpsd.dat' in raw:
            q = tmp_path / ('inpsd_' + os.path.basename(f))
            q.write_text(raw['inpsd.dat'])
            dd = asdio.read_inpsd(str(q))
            assert dd['unserved'] == [] and dd['ignored'] == [], (f, dd['unserved'], dd['ignored'])


BPULSE_FILES = {
    1: 'exp pulse\n0.0 0.0 30.0\n1.0e-16\n1\n6\n0.0\n4.0e-15\n1.0e-14\n1.4e-14\n0.01\n1.0\n',
    2: 'gaussian\n10.0 0.0 20.0\n1.0e-16\n2\n6\n0.0\n6.0e-15\n2.0e-15\n0.0\n0.0\n1.5\n',
    3: 'polexp\n0.0 25.0 0.0\n1.0e-16\n1\n6\n0.0\n5.0e-15\n2.0\n0.0\n0.0\n1.0\n',
    4: 'square\n0.0 0.0 -40.0\n1.0e-16\n3\n6\n0.0\n3.0e-15\n9.0e-15\n0.0\n0.0\n1.0\n',
}


@pytest.mark.parametrize('kind', [1, 2, 3, 4])
def test_field_pulse_schedule_equals_the_oracle_restatement(kind, tmp_path):
    """do_bpulse 1-4 (fieldpulse.f90:34-119, 178-212; evaluation times sd_driver.f90:389-393, 770-779): the product's schedule
    (uppasd_b200/fields.py) against the oracle's step-by-step restatement, and the shapes against their definitions"""
    from uppasd_b200 import fields
    p = tmp_path / 'bpulsefile'
    p.write_text(BPULSE_FILES[kind])
    P = fields.read_bpulse(str(p), kind)
    dt, rstep, nstep = 1.0e-16, 7, 160
    tf = fields.bpulse_schedule(kind, P, dt, rstep, nstep)
    B = orc.bpulse_setup(kind, P['b0'], P['step'], P['par'][:6])
    assert abs(B['ba'] - P['ba']) <= 1e-15 * abs(P['ba']) and abs(B['bb'] - P['bb']) <= 1e-15 * abs(P['bb'])
    field, scount = orc.bpulse_field(B, dt * rstep), 1
    for s in range(nstep):
        assert np.abs(tf[:, s] - field).max() <= 1e-13 * max(1.0, np.abs(field).max()), (kind, s)
        if scount == B['step']:
            field, scount = orc.bpulse_field(B, dt * (rstep + 1 + s)), 1
        else:
            scount += 1
    amp = np.abs(tf).max(axis=0) / np.abs(np.array(P['b0'])).max()
    if kind == 4:                       # square: 0 before par(2), par(6) until par(3), 0 after; held for bpulse_step steps
        assert set(np.round(amp, 12)) == {0.0, 1.0} and amp[0] == 0.0 and amp[-1] == 0.0
    if kind == 1:                       # plateau at par(6), exponential head from par(5) at par(1)
        assert abs(amp.max() - 1.0) < 1e-12 and amp[0] < 0.1
    if kind == 2:                       # Gaussian centred at par(2) with width par(3)
        assert abs(amp.max() - 1.5) < 0.01 and abs(np.argmax(amp) + rstep - 60) <= 2
    pin = tmp_path / 'inpsd.dat'
    pin.write_text('simid x\ndo_bpulse %d\nbpulsefile ./bpulsefile\n' % kind)
    d = asdio.read_inpsd(str(pin))
    assert d['do_bpulse'] == kind and d['unserved'] == [] and d['bpulsefile'].endswith('bpulsefile')
    pin.write_text('simid x\ndo_bpulse 5\n')
    assert asdio.read_inpsd(str(pin))['unserved'] == [('do_bpulse', '5')]
