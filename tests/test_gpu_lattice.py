"""On-device table construction must be bit-identical to the reference's setup_nm + mount (via the oracle)."""
import numpy as np
import pytest

from oracle import inputs, orc
from util import load_golden

pytestmark = pytest.mark.gpu


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
]


@pytest.mark.parametrize('name,over', CASES)
def test_device_tables_bit_exact(name, over):
    from uppasd_b200 import host, lattice
    fx, _, _ = load_golden(name)
    args, S = _build(fx, over)
    inp = args[0]
    kinds = [(0, 'exchange', args[6], 1, 1, inp['sym'], True)]
    if args[7] is not None:
        kinds.append((1, 'dm', args[7], 3, 1, 0, False))
    e = host.Engine()
    e.set_system(S['Natom'], 1, S['nHam'], S['aHam'])
    for kind, key, mk, ncomp, lexp, sym, typed in kinds:
        nn, red, xc, nntype = mk(S)
        ns, ca, cs, sh = lattice.stencil(inp['cell'], S['bas'], S['atype_inp'], nn, red, sym, nntype if typed else None)
        cp = lattice.couplings(ns, ca, sh, S['atype_inp'], xc, S['ammom_inp'], orc.CONST['mry'], orc.CONST['mub'], lexp)
        e.build_lattice_table(kind, S['NA'], inp['ncell'], inp['bc'], ns, ca, cs, cp)
        lst, size, coup = e.get_table(kind)
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
