"""Random alloys (do_ralloy 1, BASELINE config 3), host side: the oracle's restatement of setup_chemicaldata /
setup_neighbour_hamiltonian's alloy branch (geometry.f90:190-329, hamiltonianinit.f90:1075-1084) against its own invariants, and
the product's occupancy / readers (uppasd_b200/alloy.py, asdio.py) against the oracle.  The reference's test tree holds no
random-alloy golden: inputs are its examples/Mappings/{RandomAlloy, FeCo/random} (tests/golden/make_fixtures.py)."""
import json
import os

import numpy as np
import pytest

from oracle import inputs, orc
from util import GOLDEN
from uppasd_b200 import alloy, asdio


def _args(name, **over):
    fx = json.load(open(os.path.join(GOLDEN, name + '.json')))
    args = list(inputs.load_alloy_fixture(fx))
    args[0] = dict(args[0], **over)
    return fx, args


@pytest.mark.parametrize('name,ncell', [('randomalloy', (6, 5, 4)), ('randomalloy', (10, 10, 10)), ('feco_random', (6, 6, 6))])
def test_occupancy_counts_and_product_equals_oracle(name, ncell):
    fx, args = _args(name, ncell=ncell)
    inp, bas, atype_inp, nch, chconc = args[:5]
    na, ncellt = bas.shape[1], ncell[0] * ncell[1] * ncell[2]
    ach = orc.setup_chemicaldata(na, ncell, nch, chconc, inp['tseed'])
    assert ach.min() >= 1                                            # fully occupied
    for ia in range(na):
        for ich in range(nch[ia]):
            assert (ach[ia::na] == ich + 1).sum() == int(np.rint(chconc[ia, ich] * ncellt))
    assert np.array_equal(alloy.occupancy(na, ncell, nch, chconc, inp['tseed']), ach)      # the product's own generator + ranking
    assert not np.array_equal(orc.setup_chemicaldata(na, ncell, nch, chconc, inp['tseed'] + 1), ach)


def test_oracle_alloy_tables_are_consistent():
    """the mounted couplings are xc(chem_i, chem_j) * 2 mRy / mu_B / m_i / m_j entry by entry, the pair energy is symmetric
    (J_ij m_i m_j = J_ji m_j m_i), moments follow the species"""
    fx, args = _args('randomalloy', ncell=(6, 6, 6))
    S = orc.build_alloy_system(*args)
    inp, bas, atype_inp, nch, chconc, ammom, aemom, landeg, ex = args
    nl, ns, nc = S['exchange']['list'], S['exchange']['listsize'], S['exchange']['coup']
    N, na = S['Natom'], S['NA']
    assert (ns == 14).all()                                          # bcc, 2 shells, sym 1: 8 + 6
    site, chem = S['anumb'], S['achem_ch']
    m = ammom[site - 1, chem - 1]
    assert np.array_equal(S['mmom'][:, 0], m) and set(np.unique(m)) == {1.8, 2.23}
    nn, red, xc, nntype = ex(S)
    fc2 = 2.0 * orc.CONST['mry'] / orc.CONST['mub']
    for i in range(0, N, 7):
        for j in range(ns[i]):
            nb = nl[j, i] - 1
            shell = 0 if j < 8 else 1
            assert nc[j, i] == xc[0, 0, shell, chem[i] - 1, chem[nb] - 1] * fc2 / m[i] / m[nb]
            back = list(nl[:ns[nb], nb]).index(i + 1)
            assert abs(nc[j, i] * m[i] * m[nb] - nc[back, nb] * m[nb] * m[i]) <= 1e-12 * abs(nc[j, i] * m[i] * m[nb])


def test_product_alloy_readers_equal_the_oracle_readers(tmp_path):
    for name in ('randomalloy', 'feco_random'):
        fx, args = _args(name)
        inp, bas, atype_inp, nch, chconc, ammom, aemom, landeg, ex = args
        d = tmp_path / name
        d.mkdir()
        for k, v in fx['raw'].items():
            (d / k).write_text(v)
        pin = asdio.read_inpsd(str(d / 'inpsd.dat'))
        assert pin['do_ralloy'] == 1 and pin['unserved'] == []
        cell = np.asarray(pin['cell'], dtype=float)
        b2, t2, n2, c2 = asdio.read_posfile_alloy(pin['posfile'], cell, pin['posfiletype'])
        assert np.array_equal(b2, bas) and np.array_equal(t2, atype_inp) and np.array_equal(n2, nch) and np.array_equal(c2, chconc)
        a2, e2, l2 = asdio.read_momfile_alloy(pin['momfile'], bas.shape[1], chconc.shape[1], pin['landeg_glob'])
        assert np.array_equal(a2, ammom) and np.array_equal(e2, aemom)
        S = dict(bas=np.asfortranarray(bas))
        nn, red, xc, nntype = ex(S)
        nn2, red2, xc2, nt2 = asdio.read_pairfile_alloy(pin['exchange'], atype_inp, chconc.shape[1], bas, cell, pin['maptype'], pin['posfiletype'])
        assert np.array_equal(nn2, nn) and np.array_equal(red2, red) and np.array_equal(xc2, xc[0]) and np.array_equal(nt2, nntype)
