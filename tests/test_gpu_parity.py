"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference goldens.

Tolerances: T=0 field and trajectory within 1e-12 relative in FP64 (BASELINE.json north_star); table
construction bit-exact; thermal / Monte Carlo runs on observables within statistical error bars.
"""
import numpy as np
import pytest

from oracle import orc
from util import load_golden

pytestmark = pytest.mark.gpu

FIX = ['kagome', 'megatest', 'feco', 'bccfe_cuda', 'cluster', 'heisstripe', 'heischainaf', 'scsurf']


def _engine(S, inp, **kw):
    from uppasd_b200 import host
    args = dict(sdealgh=inp['sdealgh'], delta_t=inp['timestep'], damping=inp['damping'], temp=0.0, mompar=inp['mompar'])
    args.update(kw)
    return host.engine_from_system(S, orc.consts(S), **args)


@pytest.mark.parametrize('name', FIX)
def test_field_parity(name):
    fx, inp, S = load_golden(name)
    if inp['initmag'] == 1:
        orc.initmag1(S, inp['tseed'])
    e = _engine(S, inp)
    beff, b1, b2, en = e.effective_field(parts=True)
    rb, r1, r2, ren = orc.effective_field(S, want_parts=True)
    scale = np.abs(rb).max()
    assert np.abs(beff - rb).max() <= 1e-12 * scale
    assert np.abs(b1 - r1).max() <= 1e-12 * scale
    assert np.abs(b2 - r2).max() <= 1e-12 * scale + 1e-300
    assert abs(en[0] - ren) <= 1e-12 * abs(ren)


@pytest.mark.parametrize('name', FIX)
@pytest.mark.parametrize('alg', [1, 5])
@pytest.mark.parametrize('resident', ['1', '0'])
def test_t0_trajectory(name, alg, resident, monkeypatch):
    """resident 1: the whole time loop in one launch with the state in shared memory (llg_resident_kernel, the default for
    systems this small); resident 0: the two stage launches per step that large systems use."""
    monkeypatch.setenv('ASD_RESIDENT', resident)
    fx, inp, S = load_golden(name)
    if inp['initmag'] == 1:
        orc.initmag1(S, inp['tseed'])
    e = _engine(S, inp, sdealgh=alg)
    st = orc.SdState(S, alg, inp['timestep'], inp['damping'])
    done = 0
    for n in (1, 9, 190):
        e.sd_steps(n, first_step=done + 1)
        for _ in range(n):
            st.step()
        done += n
        emom, emomM, mmom = e.get_moments()
        assert np.abs(emom - st.emom).max() <= 1e-12, (name, alg, done)
        assert np.abs(emomM - st.emomM).max() <= 1e-12 * np.abs(st.emomM).max()
    # norm is conserved by both schemes
    assert np.abs(np.sqrt((emom ** 2).sum(axis=0)) - 1.0).max() < 1e-12


def test_kagome_golden_on_gpu():
    fx, inp, S = load_golden('kagome')
    e = _engine(S, inp)
    e.sd_steps(2400)
    emom, emomM, mmom = e.get_moments()
    exp = fx['expected']['trajectory']['2400']
    for a, b in zip(list(emom[:, 1, 0]) + [mmom[1, 0]], exp):
        assert abs(a - b) <= 1e-8
    e.sd_steps(13000 - 2400, first_step=2401)
    msum = e.measure()
    av = msum[:, 0] / S['Natom']
    got = list(av) + [float(np.sqrt((av ** 2).sum()))]
    for a, b in zip(got, fx['expected']['averages']['13000']):
        assert abs(a - b) <= 1e-8, (got, fx['expected'])


def test_megatest_golden_on_gpu():
    fx, inp, S = load_golden('megatest')
    e = _engine(S, inp)
    e.sd_steps(11000)
    emom, emomM, mmom = e.get_moments()
    exp = fx['expected']
    av = emomM[:, :, 0].sum(axis=1) / S['Natom']
    got = list(av) + [float(np.sqrt((av ** 2).sum()))]
    for a, b in zip(got, exp['averages']['11000']):
        assert abs(a - b) <= 1e-8
    for a, b in zip(emom[:, exp['moment']['atom'] - 1, 0], exp['moment']['11000']):
        assert abs(a - b) <= 1e-8


def test_legacy_boundary_kagome():
    """Drives the drop-in symbols the way FortranData_Initiate + sd_mphaseCUDA do (chelper.f90:166-186)."""
    from uppasd_b200 import host
    fx, inp, S = load_golden('kagome')
    fh = host.FortranHost(S, orc.CONST, sdealgh=1, nstep=13001, delta_t=inp['timestep'], damping=inp['damping'],
                          avrg_step=inp['avrg_step']).run()
    for a, b in zip(fh.averages[13000], fx['expected']['averages']['13000']):
        assert abs(a - b) <= 1e-8
    assert fh.flushed_at == 13002
    # final state written back to the host arrays (restart file is written from them, uppasd.f90:344-350)
    st = orc.sd_run(S, inp, nstep=13001)['state']
    assert np.abs(fh.arr['emom'] - st.emom).max() <= 1e-10
    assert np.allclose(fh.arr['mmomi'], 1.0 / fh.arr['mmom'])


def test_ensembles_and_external_field_array():
    fx, inp, S = load_golden('megatest')
    # 3 ensembles with different starting states and a non-uniform external field array
    from oracle import inputs
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], mensemble=3)
    S = orc.build_system(*args)
    rng = np.random.default_rng(7)
    N, M = S['Natom'], 3
    e0 = rng.normal(size=(3, N, M))
    e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    S['external_field'] = np.asfortranarray(rng.normal(size=(3, N, M)) * 0.1)
    e = _engine(S, inp)
    st = orc.SdState(S, 1, inp['timestep'], inp['damping'])
    e.sd_steps(50)
    for _ in range(50):
        st.step()
    emom, _, _ = e.get_moments()
    assert np.abs(emom - st.emom).max() <= 1e-12


def test_anisotropy_bq_mompar_terms():
    """Uniaxial+cubic (taniso 7/1/2), biquadratic and mompar paths against the oracle on a perturbed megaTest."""
    fx, inp, S = load_golden('megatest')
    N = S['Natom']
    rng = np.random.default_rng(11)
    ta = np.array([1, 2, 7, 0] * (N // 4), dtype=np.int32)
    ea = rng.normal(size=(3, N)); ea /= np.sqrt((ea ** 2).sum(axis=0))
    S['aniso'] = dict(taniso=ta, eaniso=np.asfortranarray(ea), kaniso=np.asfortranarray(rng.normal(size=(2, N)) * 0.5),
                      sb=rng.uniform(0.1, 1.0, size=N))
    # biquadratic table: reuse the exchange lists with their own couplings
    ex = S['exchange']
    S['bq'] = dict(list=ex['list'], listsize=ex['listsize'], coup=np.asfortranarray(ex['coup'] * 0.03), z=ex['z'])
    e0 = rng.normal(size=(3, N, 1)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    for mompar in (0, 1, 2):
        e = _engine(S, inp, mompar=mompar)
        beff, en = e.effective_field()
        rb, ren = orc.effective_field(S)
        assert np.abs(beff - rb).max() <= 1e-12 * np.abs(rb).max()
        assert abs(en[0] - ren) <= 1e-11 * abs(ren)
        for alg in (1, 5):
            e = _engine(S, inp, mompar=mompar, sdealgh=alg)
            st = orc.SdState(S, alg, inp['timestep'], inp['damping'], mompar=mompar)
            e.sd_steps(40)
            for _ in range(40):
                st.step()
            emom, emomM, mmom = e.get_moments()
            assert np.abs(emom - st.emom).max() <= 1e-12
            assert np.abs(mmom - st.mmom).max() <= 1e-12 * st.mmom.max()


def test_full_hamiltonian_matches_reduced():
    """do_reduced N (NH = Natom, per-atom couplings) must give the same dynamics as do_reduced Y."""
    from oracle import inputs
    fx, inp, S = load_golden('bccfe_cuda')
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], do_reduced='N')
    Sf = orc.build_system(*args)
    assert Sf['nHam'] == Sf['Natom']
    e1, e2 = _engine(S, inp, sdealgh=1), _engine(Sf, inp, sdealgh=1)
    e1.sd_steps(100); e2.sd_steps(100)
    a, _, _ = e1.get_moments(); b, _, _ = e2.get_moments()
    assert np.abs(a - b).max() <= 1e-13
    st = orc.SdState(Sf, 1, inp['timestep'], inp['damping'])
    for _ in range(100):
        st.step()
    assert np.abs(b - st.emom).max() <= 1e-12


def test_errors_are_loud():
    from uppasd_b200 import host
    fx, inp, S = load_golden('kagome')
    e = host.Engine()
    with pytest.raises(host.AsdError):
        e.commit()
    e.set_system(S['Natom'], 1, S['nHam'], S['aHam'])
    with pytest.raises(host.AsdError):
        e.set_llg(2, 1e-16)            # Heun is not on this path
    bad = S['exchange']['list'].copy(order='F'); bad[0, 5] = 0
    e.set_exchange(bad, S['exchange']['listsize'], S['exchange']['coup'])
    with pytest.raises(host.AsdError):
        e.commit()                     # reduced Hamiltonian with a missing neighbour


def test_legacy_style_mc_and_initial_phase_entries():
    """cudamcsim_evolve_ / cudamdsim_initialphase_ (new F77-style siblings of the legacy symbols) drive the same engine
    as the explicit API: same seed -> identical states, written back into the host's Fortran arrays."""
    from uppasd_b200 import host
    fx, inp, S = load_golden('bccfe_cuda')
    fh = host.FortranHost(S, orc.CONST, sdealgh=1, nstep=10, delta_t=1e-16, damping=0.5, gpu_rng_seed=77).initiate()
    e = host.engine_from_system(S, orc.CONST, sdealgh=1, delta_t=1e-16, damping=0.5, temp=0.0, seed=77)
    for mode in ('M', 'H'):
        fh.mc_evolve(mode, 30, 400.0, first_sweep=1)
        e.mc_sweeps(mode, 30, 400.0, first_sweep=1)
        emom, emomM, mmom = e.get_moments()
        assert np.array_equal(fh.arr['emom'], emom) and np.array_equal(fh.arr['emomM'], emomM)
    # a thermal SD initial phase with its own step, temperature and damping, then back to the measurement parameters
    fh.initial_phase(40, 300.0, 5e-16, 0.3, 5, first_step=1)
    e.set_llg(5, 5e-16, landeg=S['Landeg'], lambda1=0.3, temp=300.0, seed=77)
    e.sd_steps(40, first_step=1)
    assert np.array_equal(fh.arr['emom'], e.get_moments()[0])
    assert np.allclose(fh.arr['mmomi'], 1.0 / fh.arr['mmom'])


def test_per_site_damping_temperature_and_lande_arrays():
    """lambda1_array(N), Temp_array(N), Landeg(N) that differ from site to site (evolution.f90:38-44) take the in-kernel
    per-site branch; T = 0 so that the comparison with the oracle is deterministic."""
    fx, inp, S = load_golden('megatest')
    N = S['Natom']
    rng = np.random.default_rng(5)
    lam = rng.uniform(0.05, 0.9, size=N)
    lg = rng.uniform(0.8, 1.2, size=N)
    from uppasd_b200 import host
    for alg in (1, 5):
        e = host.engine_from_system(S, orc.CONST, sdealgh=alg, delta_t=inp['timestep'], damping=0.1, temp=0.0)
        e.set_llg(alg, inp['timestep'], landeg=lg, lambda1=lam, temp=np.zeros(N))
        S2 = dict(S, Landeg=lg)
        st = orc.SdState(S2, alg, inp['timestep'], lam)
        e.sd_steps(60)
        for _ in range(60):
            st.step()
        assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12, alg


def test_fused_moment_sum_equals_the_standalone_reduction():
    """asd_measure right after asd_sd_steps adds the per-tile partial sums that the corrector launch left behind;
    asking for the energy as well takes the stand-alone reduction over the spins.  Same numbers, and both equal the
    host-side sum of emomM."""
    from uppasd_b200 import host
    _, inp, S = load_golden('megatest')
    from oracle import inputs
    import json, os
    from util import GOLDEN
    fx = json.load(open(os.path.join(GOLDEN, 'megatest.json')))
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], mensemble=3)
    S = orc.build_system(*args)
    e = host.engine_from_system(S, orc.CONST, sdealgh=1, delta_t=inp['timestep'], damping=0.3, temp=200.0, seed=4)
    for n in (1, 7):
        e.sd_steps(n)
        fused = e.measure()
        alone, _ = e.measure(energy=True)
        _, emomM, _ = e.get_moments()
        assert np.allclose(fused, alone, rtol=1e-13, atol=1e-10)
        assert np.allclose(fused, emomM.sum(axis=1), rtol=1e-13, atol=1e-10)
    e.mc_sweeps('M', 2, 200.0)
    after_mc = e.measure()
    assert np.allclose(after_mc, e.get_moments()[1].sum(axis=1), rtol=1e-13, atol=1e-10)   # stale partials are not reused


def test_ragged_lists_isolated_atoms_and_odd_sizes():
    """Per-atom Hamiltonian (do_reduced N) whose neighbour lists have every length from 0 (an isolated moment that only
    feels the external field) to z, on a system whose size is not a multiple of the warp or tile size; three
    ensembles.  Field and both integrators against the oracle."""
    from oracle import inputs
    from uppasd_b200 import host
    import json, os
    from util import GOLDEN
    fx = json.load(open(os.path.join(GOLDEN, 'megatest.json')))
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], do_reduced='N', ncell=(3, 3, 5), mensemble=3, hfield=(0.3, -0.2, 0.5))
    S = orc.build_system(*args)
    N = S['Natom']
    assert N % 32 != 0 and S['nHam'] == N
    rng = np.random.default_rng(12)
    ex = S['exchange']
    size = ex['listsize'].copy()
    for i in range(N):
        size[i] = rng.integers(0, size[i] + 1)
    size[5] = 0
    lst = ex['list'].copy(order='F')
    coup = ex['coup'].copy(order='F')
    for i in range(N):
        lst[size[i]:, i] = 0
        coup[size[i]:, i] = 0.0
    S['exchange'] = dict(ex, list=lst, listsize=size, coup=coup)
    e0 = rng.normal(size=(3, N, 3)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    for alg in (1, 5):
        e = host.engine_from_system(S, orc.CONST, sdealgh=alg, delta_t=1e-16, damping=0.4, temp=0.0)
        beff, en = e.effective_field()
        rb, ren = orc.effective_field(S)
        assert np.abs(beff - rb).max() <= 1e-12 * np.abs(rb).max()
        assert np.abs(beff[:, 5, :] - np.array([0.3, -0.2, 0.5])[:, None]).max() == 0.0      # the isolated moment
        st = orc.SdState(S, alg, 1e-16, 0.4)
        e.sd_steps(80)
        for _ in range(80):
            st.step()
        assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12, alg


@pytest.mark.parametrize('solver', [1, 5])
def test_solvers_golden_random_start(solver):
    """tests/Solvers (regulartests.yaml:349-385): 100-spin chain from the reference's random start (Initmag 1), damping 1,
    dt 1e-15, 8000 steps: the GPU path reproduces the reference's printed averages for the midpoint and the Depondt
    solver at the reference's own tolerance (1e-8), and the oracle's final state to 1e-10."""
    from uppasd_b200 import host
    fx, inp, S = load_golden('solvers')
    orc.initmag1(S, inp['tseed'])
    e = host.engine_from_system(S, orc.CONST, sdealgh=solver, delta_t=inp['timestep'], damping=inp['damping'], temp=0.0)
    e.sd_steps(8000)
    m = e.measure()[:, 0] / S['Natom']
    got = list(m) + [float(np.sqrt((m ** 2).sum()))]
    for a, b in zip(got, fx['expected']['averages'][str(solver)]['8000']):
        assert abs(a - b) <= 1e-8, (solver, got)
    st = orc.SdState(S, solver, inp['timestep'], inp['damping'])
    for _ in range(8000):
        st.step()
    assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-10


def test_monte_carlo_refuses_tensor_exchange():
    """do_jtensor 1 is served by the LLG path only; a Monte Carlo sweep must fail loudly, not run with the wrong couplings"""
    from uppasd_b200 import host
    fx, inp, S = load_golden('kagome_cuda')
    e = _engine(S, inp)
    with pytest.raises(host.AsdError):
        e.mc_sweeps('M', 1, 10.0)


def test_tensor_exchange_field_trajectory_and_golden():
    """Tensorial exchange (do_jtensor 1; hamiltonianactions.f90:499-542) on tests/kagome_cuda: field and both solvers'
    trajectories against the oracle to 1e-12, the reference's printed averages @1300 (cudatests.yaml:1-23, 1e-8), and the
    same run through the legacy boundary (fd.j_tensor)."""
    from uppasd_b200 import host
    fx, inp, S = load_golden('kagome_cuda')
    orc.initmag1(S, inp['tseed'])
    e = _engine(S, inp)
    beff, en = e.effective_field()
    rb, ren = orc.effective_field(S)
    assert np.abs(beff - rb).max() <= 1e-12 * np.abs(rb).max()
    assert abs(en[0] - ren) <= 1e-12 * abs(ren)
    for alg in (1, 5):
        e = _engine(S, inp, sdealgh=alg)
        st = orc.SdState(S, alg, inp['timestep'], inp['damping'])
        e.sd_steps(100)
        for _ in range(100):
            st.step()
        assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12, alg
    e.sd_steps(1200, first_step=101)
    m = e.measure()[:, 0] / S['Natom']
    got = list(m) + [float(np.sqrt((m ** 2).sum()))]
    for a, b in zip(got, fx['expected']['averages']['1300']):
        assert abs(a - b) <= 1e-8, got
    fh = host.FortranHost(S, orc.CONST, sdealgh=5, nstep=1301, delta_t=inp['timestep'], damping=inp['damping'],
                          avrg_step=inp['avrg_step']).run()
    for a, b in zip(fh.averages[1300], fx['expected']['averages']['1300']):
        assert abs(a - b) <= 1e-8


@pytest.mark.parametrize('name', FIX + ['kagome_cuda'])
def test_energy_terms_parity(name):
    """asd_energy_terms against the oracle's calc_energy restatement (energy.f90:181-398) on every fixture: exchange / pair,
    anisotropy, DM, biquadratic, Zeeman per atom in mRy, 1e-12 relative to the largest term."""
    fx, inp, S = load_golden(name)
    if inp['initmag'] == 1:
        orc.initmag1(S, inp['tseed'])
    e = _engine(S, inp)
    e.sd_steps(25)
    emomM = e.get_moments()[1]
    got = e.energy_terms()
    ref = orc.energy_terms(S, emomM)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), (got, ref)
    assert abs(ref[0, 0]) > 0


def _gpu_rows(e, S, marks):
    out, done = {}, 0
    for it in sorted(marks):
        if it > done:
            e.sd_steps(it - done, first_step=done + 1)
            done = it
        m = e.measure()[:, 0] / S['Natom']
        out[it] = (list(m) + [float(np.sqrt((m ** 2).sum()))], e.energy_terms()[:, 0])
    return out


def test_cluster_golden_on_gpu():
    """tests/Cluster (regulartests.yaml:182-205, 1e-8): BQ + type-7 anisotropy + Depondt through the two initial phases
    (set_llg between phases, as sd_iphase does) and the undamped measurement phase."""
    fx, inp, S = load_golden('cluster')
    orc.initmag1(S, inp['tseed'])
    e = _engine(S, inp, sdealgh=5, delta_t=fx['ip_phases'][0]['timestep'], damping=fx['ip_phases'][0]['damping'])
    e.sd_steps(fx['ip_phases'][0]['nstep'])
    ph = fx['ip_phases'][1]
    e.set_llg(5, ph['timestep'], landeg=S['Landeg'], lambda1=ph['damping'], temp=0.0)
    e.sd_steps(ph['nstep'])
    e.set_llg(5, inp['timestep'], landeg=S['Landeg'], lambda1=inp['damping'], temp=0.0)
    r = _gpu_rows(e, S, {15000, 25000})
    exp = fx['expected']
    for a, b in zip(r[25000][0], exp['averages']['25000']):
        assert abs(a - b) <= 1e-8, (a, b)
    t, en = r[15000][1], exp['totenergy']['15000']
    for a, b in ((t.sum(), en['tot']), (t[0], en['exc']), (t[1], en['ani']), (t[3], en['bq'])):
        assert abs(a - b) <= 1e-8, (a, b)


@pytest.mark.parametrize('name', ['heisstripe', 'heischainaf', 'scsurf'])
def test_more_reference_goldens_on_gpu(name):
    """tests/HeisStripe, tests/HeisChainAF, tests/SCsurf (regulartests.yaml:54-129,158-170, 1e-8): printed averages and
    energy columns from the GPU path."""
    fx, inp, S = load_golden(name)
    orc.initmag1(S, inp['tseed'])
    e = _engine(S, inp)
    exp = fx['expected']
    marks = {int(k) for k in exp['averages']} | {int(k) for k in exp.get('totenergy', {})}
    r = _gpu_rows(e, S, marks)
    for k, v in exp['averages'].items():
        for a, b in zip(r[int(k)][0], v):
            assert abs(a - b) <= 1e-8, (name, k, a, b)
    for k, v in exp.get('totenergy', {}).items():
        t = r[int(k)][1]
        got = {'tot': t.sum(), 'exc': t[0], 'ani': t[1], 'dm': t[2], 'bq': t[3]}
        for key, b in v.items():
            assert abs(got[key] - b) <= 1e-8, (name, k, key, got[key], b)


@pytest.mark.parametrize('alg', [1, 5])
def test_resident_kernel_is_bit_identical_to_the_stage_launches(alg, monkeypatch):
    """Thermal run (300 K, 3 ensembles, DM + anisotropy, external field) of the kagome fixture: one resident launch for 200
    steps, the same 200 steps in uneven batches, and the two-launch-per-step path give the same bits (same summation order,
    same noise counters), and the launch counts differ as designed."""
    fx, inp, S0 = load_golden('kagome')
    from uppasd_b200 import host
    S = dict(S0, Mensemble=3)
    for k in ('emom', 'emomM', 'external_field'):
        S[k] = np.asfortranarray(np.repeat(S0[k], 3, axis=2))
    for k in ('mmom', 'mmom0', 'mmomi'):
        S[k] = np.asfortranarray(np.repeat(S0[k], 3, axis=1))
    S['external_field'][1] += 3.0
    out = {}
    for tag, res, batches in (('resident', '1', (200,)), ('batched', '1', (1, 7, 192)), ('stages', '0', (200,))):
        monkeypatch.setenv('ASD_RESIDENT', res)
        e = host.engine_from_system(S, orc.CONST, sdealgh=alg, delta_t=inp['timestep'], damping=0.2, temp=300.0, seed=5)
        n0, done = e.launch_count(), 0
        for n in batches:
            e.sd_steps(n, first_step=done + 1)
            done += n
        out[tag] = (e.get_moments()[0].copy(), e.launch_count() - n0)
    assert np.array_equal(out['resident'][0], out['batched'][0])
    assert np.array_equal(out['resident'][0], out['stages'][0])
    # launch counts (get_moments adds a few conversion launches to each): 1 / 3 resident launches against 400 stage launches
    assert out['batched'][1] - out['resident'][1] == 2 and out['stages'][1] - out['resident'][1] == 399, {k: v[1] for k, v in out.items()}
    assert np.abs(out['resident'][0] - S['emom']).max() > 1e-3


def test_observables_on_host_table_layouts():
    """asd_measure_sublattice and asd_skyrmion_number on layouts built from HOST tables (three sublattices of the kagome
    fixture, two ensembles, after a few steps): against numpy / the oracle's pontryagin_tri, in the LLG layout and after
    the state moved to the Monte Carlo (colour-major) layout."""
    from uppasd_b200 import host
    fx, inp, S0 = load_golden('kagome')
    S = dict(S0, Mensemble=2)
    for k in ('emom', 'emomM', 'external_field'):
        S[k] = np.asfortranarray(np.repeat(S0[k], 2, axis=2))
    for k in ('mmom', 'mmom0', 'mmomi'):
        S[k] = np.asfortranarray(np.repeat(S0[k], 2, axis=1))
    e = host.engine_from_system(S, orc.CONST, sdealgh=1, delta_t=inp['timestep'], damping=0.1, temp=50.0, seed=3)
    e.sd_steps(60)
    n1, n2, n3 = S['ncell']
    simp = orc.delaunay_tri_tri(n1, n2, n3, S['NA'])
    e.set_triangulation(simp)
    for phase in ('llg', 'mc'):
        if phase == 'mc':
            e.mc_sweeps('H', 2, 20.0)
        emom, emomM, _ = e.get_moments()
        ms = e.measure_sublattice(S['NA'])
        ref = np.stack([emomM[:, c::S['NA'], :].sum(axis=1) for c in range(S['NA'])], axis=1)
        assert np.abs(ms - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), phase
        q = e.skyrmion_number()
        assert np.abs(q - orc.pontryagin_tri(emom, simp)[1]).max() <= 1e-12 * max(1.0, np.abs(q).max()), phase
    with pytest.raises(host.AsdError):
        e.measure_sublattice(5)                    # does not divide Natom


@pytest.mark.parametrize('alg', [1, 5])
@pytest.mark.parametrize('resident', ['1', '0'])
def test_fixed_moment_list(alg, resident, monkeypatch):
    """Nred / red_atom_list of evolve_first (evolution.f90:38-44): every fifth atom of the kagome fixture is left out of the
    list.  Frozen atoms keep their moment bit for bit, still act on their neighbours, and the trajectory equals the oracle's
    (which restores the rows of the atoms the reference's loops never visit) to 1e-12; clearing the list restores the
    ordinary run."""
    monkeypatch.setenv('ASD_RESIDENT', resident)
    fx, inp, S = load_golden('kagome')
    N = S['Natom']
    fro = np.arange(N) % 5 == 0
    red = np.arange(1, N + 1)[~fro]
    e = _engine(S, inp, sdealgh=alg)
    e.set_evolving_atoms(red)
    st = orc.SdState(S, alg, inp['timestep'], inp['damping'], red_atom_list=red)
    e.sd_steps(150)
    for _ in range(150):
        st.step()
    emom = e.get_moments()[0]
    assert np.array_equal(emom[:, fro, 0], S['emom'][:, fro, 0])
    assert np.abs(emom - st.emom).max() <= 1e-12
    assert np.abs(emom[:, ~fro, 0] - S['emom'][:, ~fro, 0]).max() > 0.1
    e.set_evolving_atoms(None)
    e.set_moments(S['emom'], S['mmom'], S['mmom0'])
    e.sd_steps(50)
    st = orc.SdState(S, alg, inp['timestep'], inp['damping'])
    for _ in range(50):
        st.step()
    assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12


def test_cluster_golden_through_the_legacy_boundary():
    """tests/Cluster the way a patched sd_iphase / sd_mphaseCUDA would run it: FortranData_Initiate with the biquadratic table
    handed over through fortrandata_setextras_, the two initial phases through cudamdsim_initialphase_, the measurement phase
    through cudamdsim_measurementphase_ with the measurement callbacks: the reference's printed averages @25000
    (regulartests.yaml:186-194, 1e-8) and the final state written back to the host arrays."""
    from uppasd_b200 import host
    fx, inp, S = load_golden('cluster')
    orc.initmag1(S, inp['tseed'])
    fh = host.FortranHost(S, orc.consts(S), sdealgh=5, nstep=25001, delta_t=inp['timestep'], damping=inp['damping'],
                          avrg_step=inp['avrg_step']).initiate()
    step = 1
    for ph in fx['ip_phases']:
        fh.initial_phase(ph['nstep'], ph['temp'], ph['timestep'], ph['damping'], 5, first_step=step)
        step += ph['nstep']
    fh.lib.cudamdsim_measurementphase_()
    for a, b in zip(fh.averages[25000], fx['expected']['averages']['25000']):
        assert abs(a - b) <= 1e-8, (fh.averages[25000], fx['expected'])
    assert np.abs(np.sqrt((fh.arr['emom'] ** 2).sum(axis=0)) - 1.0).max() < 1e-10


def test_resident_kernel_with_the_largest_cluster(monkeypatch):
    """2000 atoms (bcc 10^3, z = 50): eight CTAs per cluster, 155 KB of shared memory each -- the upper end of the resident
    kernel's range.  Bit-identical to the stage launches at 300 K, and to the oracle to 1e-12 at T = 0.  (If the device could not
    co-schedule such a cluster the engine falls back to the stage launches; the results are the same either way.)"""
    import json
    import os
    from oracle import inputs
    from util import GOLDEN
    from uppasd_b200 import host
    fx = json.load(open(os.path.join(GOLDEN, 'bccfe_cuda.json')))
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], ncell=(10, 10, 10), mensemble=2)
    S = orc.build_system(*args)
    inp = args[0]
    rng = np.random.default_rng(8)
    e0 = rng.normal(size=(3, S['Natom'], 2)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    out = {}
    for res in ('1', '0'):
        monkeypatch.setenv('ASD_RESIDENT', res)
        e = host.engine_from_system(S, orc.CONST, sdealgh=1, delta_t=inp['timestep'], damping=0.3, temp=300.0, seed=2)
        e.sd_steps(25)
        out[res] = e.get_moments()[0]
    assert np.array_equal(out['1'], out['0'])
    monkeypatch.setenv('ASD_RESIDENT', '1')
    e = host.engine_from_system(S, orc.CONST, sdealgh=5, delta_t=inp['timestep'], damping=0.3, temp=0.0)
    st = orc.SdState(S, 5, inp['timestep'], 0.3)
    e.sd_steps(30)
    for _ in range(30):
        st.step()
    assert np.abs(e.get_moments()[0] - st.emom).max() <= 1e-12


@pytest.mark.parametrize('alg', [1, 5])
@pytest.mark.parametrize('path', ['resident', 'stage', 'lattice'])
def test_spin_transfer_torque_field(alg, path, monkeypatch):
    """btorque /= 0 (stt /= 'N'): midpoint adds -btorque to a1 in both half steps (midpoint.f90:86-97,134-136, :242-252),
    Depondt adds +btorque to bdup (depondt.f90:100-113,152-154, :255-262).  Every launch path: the resident small-system
    kernel, the one-atom-per-thread stage launches, and the run kernel of a device-built lattice (per-slot arrays)."""
    from util import fixture_args, lattice_engine
    monkeypatch.setenv('ASD_RESIDENT', '1' if path == 'resident' else '0')
    if path == 'lattice':
        args = fixture_args('bccfe_cuda', mens=2, ncell=(64, 4, 4), do_reduced='Y')
    else:
        args = fixture_args('kagome', mens=2)
    S = orc.build_system(*args)
    inp = args[0]
    rng = np.random.default_rng(12)
    e0 = rng.normal(size=S['emom'].shape); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0); S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    beff, _ = orc.effective_field(S)
    bt = np.asfortranarray(rng.normal(size=S['emom'].shape) * 0.3 * np.abs(beff).max())     # a torque comparable with the field
    from uppasd_b200 import host
    if path == 'lattice':
        # lattice_engine commits: the torque must be set before (asd_set_torque invalidates the commit)
        from uppasd_b200 import lattice
        e = host.Engine()
        c = orc.consts(S)
        e.set_constants(c['gama'], c['k_bolt'], c['mub'], c['mry'])
        e.set_system(S['Natom'], 2, S['nHam'], S['aHam'])
        nn, red, xc, nntype = args[6](S)
        ns, ca, cs, sh = lattice.stencil(inp['cell'], S['bas'], S['atype_inp'], nn, red, inp['sym'], nntype)
        cp = lattice.couplings(ns, ca, sh, S['atype_inp'], xc, S['ammom_inp'], c['mry'], c['mub'])
        e.build_lattice_table(0, S['NA'], inp['ncell'], inp['bc'], ns, ca, cs, cp)
        e.set_torque(bt)
        e.set_llg(alg, inp['timestep'], landeg=S['Landeg'], lambda1=0.2, temp=0.0)
        e.set_moments(S['emom'], S['mmom'])
        e.commit()
        assert e.layout_info()['runs'] == 4
    else:
        e = host.Engine()
        c = orc.consts(S)
        e.set_constants(c['gama'], c['k_bolt'], c['mub'], c['mry'])
        e.set_system(S['Natom'], 2, S['nHam'], S['aHam'])
        e.set_exchange(S['exchange']['list'], S['exchange']['listsize'], S['exchange']['coup'])
        e.set_dm(S['dm']['list'], S['dm']['listsize'], S['dm']['coup'])
        e.set_external_field(S['external_field'])
        e.set_torque(bt)
        e.set_llg(alg, inp['timestep'], landeg=S['Landeg'], lambda1=0.2, temp=0.0)
        e.set_moments(S['emom'], S['mmom'], S['mmom0'])
        e.commit()
    st = orc.SdState(S, alg, inp['timestep'], 0.2, btorque=bt)
    st0 = orc.SdState(S, alg, inp['timestep'], 0.2)
    e.sd_steps(50)
    for _ in range(50):
        st.step(); st0.step()
    emom = e.get_moments()[0]
    assert np.abs(st.emom - st0.emom).max() > 1e-6          # the torque matters at this amplitude
    assert np.abs(emom - st.emom).max() <= 1e-12, (alg, path, np.abs(emom - st.emom).max())


def test_pyasd_entry_points():
    """relax_ / get_emom_ / put_emom_ / get_beff_ / get_energy_ (source/pyasd.f90:255-384, 505-517), called by reference like
    the Python package uppasd calls them, on the engine behind the legacy boundary: the field and the energy per atom against
    the oracle's effective_field, a T = 0 relax_ against the oracle's sd_minimal loop (midpoint, the module's time step, the
    damping of the call; itimestep unused as in the reference), put_emom_ -> get_emom_ round trip, Monte Carlo relax_ against
    the explicit API with the same seed."""
    from uppasd_b200 import host
    fx, inp, S = load_golden('kagome')
    N, M = S['Natom'], S['Mensemble']
    fh = host.FortranHost(S, orc.consts(S), sdealgh=5, nstep=10, delta_t=inp['timestep'], damping=inp['damping'], gpu_rng_seed=5).initiate()
    rb, ren = orc.effective_field(S)
    beff = fh.get_beff()
    assert np.abs(beff - rb).max() <= 1e-12 * np.abs(rb).max()
    assert np.array_equal(fh.arr['beff'], beff)
    assert abs(fh.get_energy() - ren / (N * M)) <= 1e-12 * abs(ren / (N * M))      # the oracle's energy: all atoms, all ensembles
    # relax_ in SD mode: solver 1 whatever SDEalgh the run uses, damping of the call, the module's delta_t
    mom = fh.relax('S', 25, 0.0, 123.0, 0.3)
    st = orc.SdState(S, 1, inp['timestep'], 0.3)
    for _ in range(25):
        st.step()
    assert np.abs(fh.arr['emom'] - st.emom).max() <= 1e-12
    assert np.abs(mom - st.emom * S['mmom'][None]).max() <= 1e-12 * np.abs(S['mmom']).max()
    assert np.array_equal(mom, fh.arr['emomM']) and np.array_equal(fh.get_emom(), fh.arr['emom'])
    # put_emom_: new directions, emomM = moments * mmom, the device state follows
    rng = np.random.default_rng(3)
    e1 = rng.normal(size=(3, N, M)); e1 /= np.sqrt((e1 ** 2).sum(axis=0))
    fh.put_emom(e1)
    assert np.array_equal(fh.get_emom(), np.asfortranarray(e1))
    assert np.array_equal(fh.arr['emomM'], e1 * fh.arr['mmom'][None]) and np.array_equal(fh.arr['emom2'], e1)
    S2 = dict(S, emom=np.asfortranarray(e1), emomM=np.asfortranarray(e1 * S['mmom'][None]))
    rb2, _ = orc.effective_field(S2)
    assert np.abs(fh.get_beff() - rb2).max() <= 1e-12 * np.abs(rb2).max()
    # relax_ in Monte Carlo mode = nsweeps of mc_evolve with the run's seed; the draw counter continues after the 25 SD steps
    e = host.engine_from_system(S2, orc.consts(S), temp=0.0, seed=5)
    for mode in ('M', 'H'):
        first = 26 if mode == 'M' else 36
        momc = fh.relax(mode, 10, 5.0, 0.0, 0.0)
        e.mc_sweeps(mode, 10, 5.0, first_sweep=first, extfield=S['external_field'][:, 0, 0])
        assert np.array_equal(momc, e.get_moments()[1]), mode


def test_legacy_measurement_phase_samples_asynchronously(monkeypatch):
    """cudamdsim_measurementphase_: sampled steps are staged on the device, land in pinned host memory on a copy stream and are
    measured by a worker thread while the time loop goes on (gpu_files/cudaMeasurement.cu:109-182, measurementQueue.cpp:63-121 in
    the reference).  Every sample must be the state of ITS step, delivered in order: identical to the blocking path
    (ASD_LEGACY_SYNC=1), also when samples are denser than the ring is deep and the host routine is slower than a step."""
    import time
    from uppasd_b200 import capi, host
    fx, inp, S = load_golden('bccfe_cuda')
    out = {}
    for tag, sync in (('async', '0'), ('sync', '1')):
        monkeypatch.setenv('ASD_LEGACY_SYNC', sync)
        fh = host.FortranHost(S, orc.CONST, sdealgh=5, nstep=60, delta_t=1e-16, damping=0.1, temp=300.0, gpu_rng_seed=3, avrg_step=2)
        seen = []
        plain = fh._measure_moment

        def slow(emomM, emom, mmom, mstep, plain=plain, seen=seen):
            time.sleep(0.002)
            e = np.ctypeslib.as_array(emom, shape=(fh.M, fh.N, 3)).copy()
            m = np.ctypeslib.as_array(mmom, shape=(fh.M, fh.N)).copy()
            eM = np.ctypeslib.as_array(emomM, shape=(fh.M, fh.N, 3))
            assert np.array_equal(eM, e * m[:, :, None])
            seen.append(mstep[0])
            plain(emomM, emom, mmom, mstep)
        fh._cbs = (fh._cbs[0], capi.CB_MEASURE(slow), fh._cbs[2], fh._cbs[3])
        before = fh.lib.asd_legacy_async_samples()
        fh.run()
        served = fh.lib.asd_legacy_async_samples() - before
        out[tag] = (dict(fh.averages), list(seen), served, fh.arr['emom'].copy())
    a, s = out['async'], out['sync']
    assert s[2] == 0 and a[2] == 30                      # 30 sampled steps went through the ring; the final sample is synchronous
    assert a[1] == s[1] == sorted(s[1]) and len(a[1]) == 31
    assert a[0].keys() == s[0].keys()
    for k in a[0]:
        assert a[0][k] == s[0][k], k                     # the very same numbers
    assert np.array_equal(a[3], s[3])


@pytest.mark.parametrize('alg', [1, 5])
@pytest.mark.parametrize('path', ['resident', 'stage', 'runs'])
def test_time_dependent_uniform_field(alg, path, monkeypatch):
    """asd_set_time_field: the global time-dependent field (pulse / microwave field of calc_external_time_fields,
    calculatefields.f90:92-185) that effective_field adds to beff2 next to external_field (hamiltonianactions.f90:241), one
    vector per step, handed to the stage kernels as a parameter from a schedule set once.  Oracle: the same steps with
    external_field + time field as its external field; steps outside the schedule see none."""
    from uppasd_b200 import host
    monkeypatch.setenv('ASD_RESIDENT', '1' if path == 'resident' else '0')
    if path == 'runs':
        import bench
        S = bench.oracle_bcc((64, 4, 4), mensemble=2)
        e, n = bench.bcc_engine((64, 4, 4), alg, 0.0, 0.5, 2, 0, 0)
        assert e.layout_info()['runs'] == 4
        dt, damp = 1e-16, 0.5
    else:
        fx, inp, S = load_golden('kagome')
        dt, damp = inp['timestep'], inp['damping']
        e = host.engine_from_system(S, orc.consts(S), sdealgh=alg, delta_t=dt, damping=damp, temp=0.0)
    M = S['Mensemble']
    nst, first = 12, 101
    rng = np.random.default_rng(5)
    tf = np.asfortranarray(rng.normal(size=(3, nst)) * 40.0)               # tesla-sized pulses: visible against the exchange field
    e.set_time_field(first + 2, tf[:, 2:10])                               # schedule covers steps first + 2 .. first + 9 only
    e.sd_steps(nst, first_step=first)
    st = orc.SdState(S, alg, dt, damp)
    ext0 = S['external_field'].copy(order='F')
    for s in range(nst):
        S['external_field'][...] = ext0
        if 2 <= s < 10:
            S['external_field'] += tf[:, s][:, None, None]
        st.step()
    S['external_field'][...] = ext0
    got = e.get_moments()[0]
    assert np.abs(got - st.emom).max() <= 1e-12
    # the field matters (a run without it ends elsewhere), and clearing the schedule removes it
    e.set_moments(S['emom'], S['mmom'])
    e.set_time_field(0, None)
    e.sd_steps(nst, first_step=first)
    assert np.abs(e.get_moments()[0] - got).max() > 1e-6
