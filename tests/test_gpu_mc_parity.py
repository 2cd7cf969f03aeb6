"""Deterministic Monte Carlo chain parity: every update path of the product against the oracle's
restatement of mc_evolve (montecarlo.f90:44-273, montecarlo_common.f90:25-79,190-200,371-422,431-865).

A colour-parallel sweep is a sequential single-site sweep in the order asd_get_mc_visit_order reports (updates inside a
colour class commute), and the draws of the counter-based generator are keyed by (atom, ensemble, sweep): the oracle is
given that visiting order (mc_evolve's iflip_a) and those very draws (asd_debug_mc_draws: the uniforms are checked bit for
bit against a numpy restatement of Philox4x32-10 here, the Gaussians to float accuracy) and must then produce the SAME
chain -- every trial move, every delta-E branch (Heisenberg, DM, biquadratic, uniaxial / cubic / type-7 anisotropy, Zeeman),
every acceptance, every heat-bath frame -- to CHAIN_TOL after several sweeps."""
import numpy as np
import pytest

from oracle import orc
from util import fixture_system, lattice_engine

pytestmark = pytest.mark.gpu

# The device contracts a*b+c into FMAs and sums the neighbours in its own order, the oracle rounds every operation
# (-ffp-contract=off) in list order: each local field differs in the last bits.  A Metropolis chain does not accumulate that
# (an accepted move IS the trial vector, which depends on the draws only), so it must agree to a few ulp.  A heat-bath
# update is a continuous function of the local field whose frame rotation divides by sin(theta) of the field direction:
# a field within 1e-3 of the z axis amplifies the last-bit difference by 1e6, and the chain carries it on (observed: up to
# 6e-10 after 5 sweeps, depending on the summation order of the kernel).  A wrong branch, sign, draw or visiting order shows
# up as O(1e-2 .. 1) in either mode.
CHAIN_TOL = {'M': 5e-12, 'H': 5e-9}


def philox4x32_10(c, k):
    """numpy restatement of Philox4x32-10 (Salmon et al. 2011): c (4, n) uint32 counters, k (2,) uint32 key"""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [x.astype(np.uint64) for x in c]
    k0, k1 = int(k[0]), int(k[1])
    for _ in range(10):
        p0, p1 = c[0] * M0, c[2] * M1
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & 0xffffffff, p1 >> 32, p1 & 0xffffffff
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0, k1 = (k0 + W0) & 0xffffffff, (k1 + W1) & 0xffffffff
    return c


def uniform4_numpy(seed, natom, mens, sweep, stream=1):
    out = np.zeros((4, natom, mens), order='F')
    key = (seed & 0xffffffff, seed >> 32)
    for k in range(mens):
        atom = np.arange(natom, dtype=np.uint64)
        s_lo, s_hi = sweep & 0xffffffff, ((sweep >> 32) ^ (stream << 24)) & 0xffffffff
        for half in range(2):
            ens = np.full(natom, k | (0x80000000 if half else 0), dtype=np.uint64)
            r = philox4x32_10([atom, ens, np.full(natom, s_lo, dtype=np.uint64), np.full(natom, s_hi, dtype=np.uint64)], key)
            out[2 * half, :, k] = (((r[0] << 32) | r[1]) >> 11).astype(np.float64) / 9007199254740992.0
            out[2 * half + 1, :, k] = (((r[2] << 32) | r[3]) >> 11).astype(np.float64) / 9007199254740992.0
    return out


def _host_engine(S, seed):
    from uppasd_b200 import host
    return host.engine_from_system(S, orc.consts(S), temp=0.0, seed=seed)


def _random_start(S, seed):
    rng = np.random.default_rng(seed)
    e0 = rng.normal(size=S['emom'].shape)
    e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])


def _chain_parity(e, S, mode, T, nsweeps, extfield=(0.0, 0.0, 0.0), dm_quirk=True, first=1):
    order = e.get_mc_visit_order()
    assert sorted(order) == list(range(1, S['Natom'] + 1))
    st = orc.McState(S, dm_quirk=dm_quirk)
    st.emom, st.emomM, st.mmom = (x.copy(order='F') for x in e.get_moments())
    e.mc_sweeps(mode, nsweeps, T, first_sweep=first, extfield=extfield)
    for s in range(first, first + nsweeps):
        u, g = e.debug_mc_draws(s)
        st.sweep(mode, T, order, u, g, extfield=extfield)
    got = e.get_moments()[0]
    err = float(np.abs(got - st.emom).max())
    moved = float(np.abs(got - S['emom']).max())
    return err, moved


def test_draws_are_the_counter_based_generator():
    """asd_debug_mc_draws = uniform4 / gauss3f of the update kernels: uniforms bit for bit Philox4x32-10 keyed by (atom, ensemble,
    sweep), Gaussians standard normal."""
    _, _, S = fixture_system('bccfe', mens=2)
    seed = 77
    e = _host_engine(S, seed)
    u, g = e.debug_mc_draws(5)
    ref = uniform4_numpy(seed ^ 0x5bd1e995, S['Natom'], 2, 5)
    assert np.array_equal(u, ref)
    u2, g2 = e.debug_mc_draws(6)
    assert not np.array_equal(u, u2)
    gg = np.concatenate([e.debug_mc_draws(s)[1].ravel() for s in range(1, 40)])
    assert abs(gg.mean()) < 5 / np.sqrt(gg.size) and abs(gg.var() - 1.0) < 0.02 and abs((gg ** 4).mean() - 3.0) < 0.1


# (fixture, overrides, temperature, Zeeman field of mc_evolve, reference DM quirk applies)
HOST_CASES = [
    ('bccfe', dict(), 300.0, (0.0, 0.0, 0.0), True),
    ('cluster', dict(), 30.0, (0.0, 0.0, 5.0), True),          # biquadratic + anisotropy type 7, BC 0 0 0
    ('scsurf', dict(), 0.5, (0.1, 0.0, 0.2), True),            # DM + anisotropy, atomic units
    ('heisstripe', dict(), 20.0, (0.0, 0.0, 0.0), True),       # uniaxial anisotropy (type 1)
    ('kagome', dict(), 5.0, (0.0, 0.0, 1.0), False),           # DM, reduced Hamiltonian with three basis atoms, |m| /= 1
    ('feco', dict(), 900.0, (0.0, 0.0, 0.0), True),            # two sublattices, z = 258
]


@pytest.mark.parametrize('mode', ['M', 'H'])
@pytest.mark.parametrize('resident', ['1', '0'])
@pytest.mark.parametrize('name,over,T,ext,quirk', HOST_CASES)
def test_chain_parity_host_tables(name, over, T, ext, quirk, resident, mode, monkeypatch):
    """colour-major layout: the resident small-system kernel and the colour launches (cooperative forms included)"""
    monkeypatch.setenv('ASD_RESIDENT', resident)
    _, inp, S = fixture_system(name, mens=2, **over)
    _random_start(S, 5)
    e = _host_engine(S, 31)
    err, moved = _chain_parity(e, S, mode, T, 6, extfield=ext, dm_quirk=quirk or bool(np.all(S['mmom'] == 1.0)))
    assert moved > 0.1
    assert err <= CHAIN_TOL[mode], (name, mode, resident, err)


def test_chain_parity_cubic_anisotropy():
    """taniso 2 (cubic) is in none of the reference's fixtures: synthetic case on the bcc Fe lattice"""
    _, inp, S = fixture_system('bccfe', mens=2)
    N = S['Natom']
    rng = np.random.default_rng(1)
    ea = rng.normal(size=(3, N)); ea /= np.sqrt((ea ** 2).sum(axis=0))
    S['aniso'] = dict(taniso=np.full(N, 2, dtype=np.int32), eaniso=np.asfortranarray(ea),
                      kaniso=np.asfortranarray(np.stack([np.full(N, 3.0), np.full(N, -1.5)])), sb=np.full(N, 0.4))
    S['aniso']['taniso'][::3] = 1
    S['aniso']['taniso'][1::3] = 7
    _random_start(S, 2)
    for mode in ('M', 'H'):
        e = _host_engine(S, 9)
        err, moved = _chain_parity(e, S, mode, 200.0, 6, extfield=(0.0, 2.0, 0.0))
        assert moved > 0.1 and err <= CHAIN_TOL[mode], (mode, err)


def _with_bq(args):
    """adds a biquadratic table on the exchange stencil (none of the reference's lattice fixtures has one with do_reduced Y)"""
    ex = args[6]
    args[8] = lambda S: (lambda t: (t[0], t[1], 0.04 * np.asarray(t[2]), None))(ex(S) if callable(ex) else ex)
    return args


def _with_aniso(S, per_row):
    """per_row: one anisotropy per basis atom (rides in the constant bank); else types 1 / 2 / 7 mixed atom by atom"""
    N, NA = S['Natom'], S['NA']
    rng = np.random.default_rng(4)
    if per_row:
        ea = np.repeat(np.array([[0.0, 0.6, 0.8]]).T, N, axis=1)
        ta = np.where(np.arange(N) % NA == 0, 7, 2).astype(np.int32)
    else:
        ea = rng.normal(size=(3, N)); ea /= np.sqrt((ea ** 2).sum(axis=0))
        ta = (np.array([1, 2, 7])[np.arange(N) % 3]).astype(np.int32)
    S['aniso'] = dict(taniso=ta, eaniso=np.asfortranarray(ea), kaniso=np.asfortranarray(np.stack([np.full(N, 3.0), np.full(N, -1.5)])),
                      sb=np.full(N, 0.4))


# device-built lattices: the block sweep (layout 2: tiles of 256 and of 1024 slots, needs do_reduced Y) and the per-colour
# tile launches (layout 1).  (fixture, overrides, T, Zeeman field, reference DM quirk applies, tile size, extras, layouts)
LATTICE_CASES = [
    ('bccfe', dict(ncell=(12, 10, 8), do_reduced='Y'), 300.0, (0.0, 0.0, 0.0), True, '256', '', (2, 1)),
    ('bccfe', dict(ncell=(64, 8, 4), do_reduced='Y'), 600.0, (0.0, 0.0, 3.0), True, '1024', '', (2, 1)),
    ('bccfe', dict(ncell=(33, 7, 5), do_reduced='Y'), 600.0, (0.0, 0.0, 0.0), True, '256', '', (2, 1)),      # odd extents: padded bricks, 3 tile colours per axis
    ('bccfe', dict(ncell=(32, 8, 8), do_reduced='Y'), 500.0, (0.0, 1.0, 0.0), True, '1024', 'bq+aniso_rows', (2, 1)),
    ('bccfe', dict(ncell=(16, 6, 6), do_reduced='Y'), 500.0, (0.0, 1.0, 0.0), True, '256', 'bq+aniso_atoms', (2, 1)),
    ('kagome', dict(ncell=(24, 12, 1)), 5.0, (0.0, 0.0, 1.0), False, '256', '', (2, 1)),                    # DM neighbours in the gather lists
    ('kagome', dict(ncell=(64, 16, 1)), 5.0, (0.0, 0.0, 1.0), False, '1024', '', (2,)),
    ('scsurf', dict(do_reduced='Y'), 0.5, (0.1, 0.0, 0.2), True, '256', '', (2, 1)),                        # DM + anisotropy, atomic units
    ('cluster', dict(), 30.0, (0.0, 0.0, 5.0), True, '256', '', (1,)),                                      # BQ + type-7 anisotropy, one cell, BC 0 0 0
    ('heisstripe', dict(), 20.0, (0.0, 0.0, 0.0), True, '256', '', (1,)),
]


@pytest.mark.parametrize('mode', ['M', 'H'])
@pytest.mark.parametrize('name,over,T,ext,quirk,ts,extra,layouts', LATTICE_CASES)
def test_chain_parity_lattice_layouts(name, over, T, ext, quirk, ts, extra, layouts, mode, monkeypatch):
    monkeypatch.setenv('ASD_RESIDENT', '0')
    monkeypatch.setenv('ASD_MC_TS', ts)
    from util import fixture_args
    args = fixture_args(name, mens=2, **over)
    if 'bq' in extra:
        args = _with_bq(args)
    S = orc.build_system(*args)
    if 'aniso' in extra:
        _with_aniso(S, 'rows' in extra)
    _random_start(S, 6)
    q = quirk or bool(np.all(S['mmom'] == 1.0))
    for layout in layouts:
        e = lattice_engine(args, S, seed=13)
        e.set_mc_layout(layout)
        lay, ncol, per = e.mc_colouring()
        assert lay == layout, (lay, layout)
        err, moved = _chain_parity(e, S, mode, T, 5, extfield=ext, dm_quirk=q)
        assert moved > 0.1
        assert err <= CHAIN_TOL[mode], (name, mode, layout, err)
        # a second batch continues the same chain (the sweep number keys the draws)
        err2, _ = _chain_parity(e, S, mode, T, 3, extfield=ext, dm_quirk=q, first=6)
        assert err2 <= CHAIN_TOL[mode], (name, mode, layout, err2)
        e.close()


def test_block_sweep_observables_and_default_choice(monkeypatch):
    """bcc Fe 64 x 32 x 32 (131 072 spins x 5 ensembles: above the size where the block sweep becomes the default): the
    layout the engine picks by itself is the block sweep with the 8-colouring of period (2, 2, 2); Metropolis and heat bath
    relax a random start towards the ordered state at 300 K and agree with each other on <|M|>."""
    import bench
    e, n = bench.bcc_engine((64, 32, 32), 1, 300.0, 0.5, 5, 0, 0)
    lay, ncol, per = e.mc_colouring()
    assert lay == 2 and ncol == 8 and per == (2, 2, 2), (lay, ncol, per)
    mags = {}
    for mode in ('M', 'H'):
        e.init_moments_tilted(0.1, bench.BCC['mom'])
        e.mc_sweeps(mode, 300, 300.0)
        acc = []
        for r in range(20):
            e.mc_sweeps(mode, 5, 300.0, first_sweep=301 + 5 * r)
            acc.append(np.sqrt(((e.measure() / n) ** 2).sum(axis=0)) / 2.23)
        mags[mode] = np.mean(acc)
        emom = e.get_moments()[0]
        assert np.abs(np.sqrt((emom ** 2).sum(axis=0)) - 1.0).max() < 1e-12
    assert 0.8 < mags['M'] < 0.995 and abs(mags['M'] - mags['H']) < 0.01, mags


@pytest.mark.parametrize('mode', ['M', 'H'])
def test_chain_parity_warp_specialised_sweep(mode, monkeypatch):
    """ASD_MC_NT=512: the warp-specialised form of the run-form block sweep (8 sweep warps + 8 draw warps, double-buffered draw
    records of 256 attempts, named barriers) runs the same chain as every other form"""
    monkeypatch.setenv('ASD_RESIDENT', '0')
    monkeypatch.setenv('ASD_MC_TS', '1024')
    monkeypatch.setenv('ASD_MC_NT', '512')
    from util import fixture_args
    args = fixture_args('bccfe', mens=2, ncell=(64, 8, 8), do_reduced='Y')
    S = orc.build_system(*args)
    _random_start(S, 8)
    e = lattice_engine(args, S, seed=17)
    e.set_mc_layout(2)
    assert e.mc_colouring()[0] == 2
    err, moved = _chain_parity(e, S, mode, 500.0, 6, extfield=(0.0, 0.5, 0.0))
    assert moved > 0.1 and err <= CHAIN_TOL[mode], (mode, err)


@pytest.mark.parametrize('mode', ['M', 'H'])
def test_chain_parity_predrawn_trial_moves(mode, monkeypatch):
    """ASD_MC_PREDRAW=1: the trial moves of a sweep drawn for the whole lattice by mc_predraw_kernel before the sweep kernel of the
    run-form block sweep (an option kept for A/B runs: measured slower than drawing inside the sweep CTAs) -- the same chain"""
    monkeypatch.setenv('ASD_RESIDENT', '0')
    monkeypatch.setenv('ASD_MC_TS', '1024')
    monkeypatch.setenv('ASD_MC_PREDRAW', '1')
    from util import fixture_args
    args = fixture_args('bccfe', mens=2, ncell=(64, 8, 8), do_reduced='Y')
    S = orc.build_system(*args)
    _random_start(S, 8)
    e = lattice_engine(args, S, seed=17)
    e.set_mc_layout(2)
    assert e.mc_colouring()[0] == 2
    err, moved = _chain_parity(e, S, mode, 500.0, 6, extfield=(0.0, 0.5, 0.0))
    assert moved > 0.1 and err <= CHAIN_TOL[mode], (mode, err)
