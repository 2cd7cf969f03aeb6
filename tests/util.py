"""Helpers shared by the test modules: load a golden fixture into an oracle system."""
import json
import os

from oracle import inputs, orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    with open(os.path.join(GOLDEN, name + '.json')) as fh:
        fx = json.load(fh)
    args = inputs.load_fixture(fx)
    inp = args[0]
    S = orc.build_system(*args)
    return fx, inp, S


def fixture_args(name, mens=None, **over):
    """the oracle's build_system arguments of a golden fixture with input overrides"""
    with open(os.path.join(GOLDEN, name + '.json')) as fh:
        fx = json.load(fh)
    args = list(inputs.load_fixture(fx))
    if mens is not None:
        over = dict(over, mensemble=mens)
    args[0] = dict(args[0], **over)
    return args


def fixture_system(name, mens=None, **over):
    """(fixture args, inp, S): the oracle system of a golden fixture with input overrides (ncell, bc, do_reduced, ...)"""
    with open(os.path.join(GOLDEN, name + '.json')) as fh:
        fx = json.load(fh)
    args = list(inputs.load_fixture(fx))
    if mens is not None:
        over = dict(over, mensemble=mens)
    args[0] = dict(args[0], **over)
    return args, args[0], orc.build_system(*args)


def lattice_engine(args, S, sdealgh=1, delta_t=1e-16, damping=0.05, temp=0.0, seed=20261017, device=-1):
    """Engine whose tables (exchange / DM / BQ) are built ON THE DEVICE from the unit-cell stencils of the fixture
    (asd_build_lattice_table: brick order, tile and run tables, block-sweep Monte Carlo); anisotropy, field and moments from S."""
    from uppasd_b200 import host, lattice
    inp = args[0]
    c = orc.consts(S)
    e = host.Engine(device)
    e.set_constants(c['gama'], c['k_bolt'], c['mub'], c['mry'])
    e.set_system(S['Natom'], S['Mensemble'], S['nHam'], S['aHam'] if S['nHam'] < S['Natom'] else None)
    kinds = [(0, args[6], 1, 1, inp['sym'], True)]
    if args[7] is not None:
        kinds.append((1, args[7], 3, 1, 0, False))
    if args[8] is not None:
        kinds.append((2, args[8], 1, 2, inp['sym'], False))
    for kind, mk, ncomp, lexp, sym, typed in kinds:
        nn, red, xc, nntype = mk(S) if callable(mk) else mk
        ns, ca, cs, sh = lattice.stencil(inp['cell'], S['bas'], S['atype_inp'], nn, red, sym, nntype if typed else None, ncell=inp['ncell'])
        cp = lattice.couplings(ns, ca, sh, S['atype_inp'], xc, S['ammom_inp'], c['mry'], c['mub'], lexp)
        e.build_lattice_table(kind, S['NA'], inp['ncell'], inp['bc'], ns, ca, cs, cp)
    if S.get('aniso') is not None:
        a = S['aniso']
        e.set_anisotropy(a['taniso'], a['eaniso'], a['kaniso'], a['sb'])
    e.set_external_field(S['external_field'])
    e.set_llg(sdealgh, delta_t, landeg=S['Landeg'], lambda1=damping, temp=temp, seed=seed)
    e.set_moments(S['emom'], S['mmom'], S['mmom0'])
    e.commit()
    return e
