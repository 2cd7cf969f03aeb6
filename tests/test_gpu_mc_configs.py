"""BASELINE.json configs 2 and 3 as parity cases: FeCo B2 two-sublattice Monte Carlo temperature points with 8
ensembles (Metropolis), and the kagome Heisenberg + DMI lattice driven by heat-bath sweeps followed by an LLG
relaxation.  Monte Carlo compares observables with the oracle's replay of mc_mphase (reference generators, random
sequential visiting order) within statistical error bars; the LLG relaxation that follows starts from the GPU's own
MC state and must match the oracle to 1e-12."""
import json
import os

import numpy as np
import pytest

from oracle import inputs, orc
from util import GOLDEN

pytestmark = pytest.mark.gpu


def _system(name, mens, **over):
    fx = json.load(open(os.path.join(GOLDEN, name + '.json')))
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], mensemble=mens, **over)
    return args[0], orc.build_system(*args)


@pytest.mark.parametrize('T', [300.0, 1100.0])
def test_feco_metropolis_temperature_points(T):
    """FeCo B2 (tests/FeCo tables: 2 sublattices, z up to 258), Mensemble = 8: <|M|> and <E> at a point well inside the
    ordered phase and at one in the critical region."""
    from uppasd_b200 import host
    inp, S = _system('feco', 8)
    e = host.engine_from_system(S, orc.CONST, temp=T, seed=21)
    e.mc_sweeps('M', 400, T)
    gm, ge = [], []
    for r in range(80):
        e.mc_sweeps('M', 5, T, first_sweep=401 + 5 * r)
        m, en = e.measure(energy=True)
        gm.append(np.sqrt(((m / S['Natom']) ** 2).sum(axis=0)))
        ge.append(en / S['Natom'])
    gm, ge = np.array(gm), np.array(ge)                # (samples, 8)
    inp4, S4 = _system('feco', 4)
    rm, re_, _ = orc.mc_run(S4, 'M', T, 800, seed=5, sample_every=5, burn=400)
    g_m, r_m = gm.mean(axis=0), rm.mean(axis=0)
    sig_m = np.sqrt(g_m.var() / len(g_m) + r_m.var() / len(r_m))
    assert abs(g_m.mean() - r_m.mean()) < 5 * sig_m + 0.03, (T, g_m.mean(), r_m.mean(), sig_m)
    g_e = ge.mean(axis=0)
    sig_e = np.sqrt(g_e.var() / len(g_e)) + abs(re_.std()) / np.sqrt(len(re_) / 10.0)
    assert abs(g_e.mean() - re_.mean()) < 5 * sig_e + 0.02 * abs(re_.mean()), (T, g_e.mean(), re_.mean(), sig_e)
    # Tc ordering: the hot point must be less magnetised than the cold one would be (sanity of the scan direction)
    emom, _, _ = e.get_moments()
    assert np.abs(np.sqrt((emom ** 2).sum(axis=0)) - 1.0).max() < 1e-12


def test_kagome_heatbath_then_llg_relaxation():
    """tests/kagome tables (Heisenberg + DMI, reduced Hamiltonian, NA = 3): heat-bath sweeps at 5 K, observables against
    the oracle; then the GPU state is relaxed with the midpoint LLG solver at T = 0 and must follow the oracle started
    from that very state (MC layout -> SD layout hand-over inside the engine)."""
    from uppasd_b200 import host
    T = 5.0
    inp, S = _system('kagome', 8)
    e = host.engine_from_system(S, orc.CONST, sdealgh=1, delta_t=inp['timestep'], damping=inp['damping'], temp=0.0, seed=8)
    e.mc_sweeps('H', 200, T)
    gm, ge = [], []
    for r in range(60):
        e.mc_sweeps('H', 5, T, first_sweep=201 + 5 * r)
        m, en = e.measure(energy=True)
        gm.append(np.sqrt(((m / S['Natom']) ** 2).sum(axis=0)))
        ge.append(en / S['Natom'])
    gm, ge = np.array(gm), np.array(ge)
    inp4, S4 = _system('kagome', 4)
    rm, re_, _ = orc.mc_run(S4, 'H', T, 500, seed=2, sample_every=5, burn=200)
    g_e = ge.mean(axis=0)
    sig_e = np.sqrt(g_e.var() / len(g_e)) + abs(re_.std()) / np.sqrt(len(re_) / 10.0)
    assert abs(g_e.mean() - re_.mean()) < 5 * sig_e + 0.02 * abs(re_.mean()), (g_e.mean(), re_.mean(), sig_e)
    g_m, r_m = gm.mean(axis=0), rm.mean(axis=0)
    sig_m = np.sqrt(g_m.var() / len(g_m) + r_m.var() / len(r_m))
    assert abs(g_m.mean() - r_m.mean()) < 5 * sig_m + 0.03, (g_m.mean(), r_m.mean(), sig_m)
    # ---- LLG relaxation from the Monte Carlo state ----
    emom, emomM, mmom = e.get_moments()
    S['emom'], S['emomM'], S['mmom'] = emom.copy(order='F'), emomM.copy(order='F'), mmom.copy(order='F')
    st = orc.SdState(S, 1, inp['timestep'], inp['damping'])
    e.sd_steps(300)
    for _ in range(300):
        st.step()
    a, _, _ = e.get_moments()
    assert np.abs(a - st.emom).max() <= 1e-12
    _, en_after = e.measure(energy=True)
    assert (en_after / S['Natom'] <= ge[-1] + 1e-9).all()      # damped T = 0 dynamics can only lower the energy


def test_mc_on_device_built_tables_matches_host_tables():
    """Monte Carlo on a lattice whose tables were built on the device (asd_build_lattice_table) = the same chain as on
    tables handed over by the host: same colouring, noise keyed by the original atom index."""
    import bench
    from uppasd_b200 import host
    e1, n = bench.bcc_engine((6, 6, 6), 1, 300.0, 0.5, 2, 0, 0)
    lst, size, coup = e1.get_table(0)
    emom, emomM, mmom = e1.get_moments()
    e2 = host.Engine(0)
    C = bench.CONST
    e2.set_constants(C['gama'], C['k_bolt'], C['mub'], C['mry'])
    e2.set_system(n, 2, 2, (np.arange(n, dtype=np.int32) % 2) + 1)
    e2.set_exchange(lst, size, coup)
    e2.set_llg(1, 1e-16, landeg=1.0, lambda1=0.5, temp=300.0, seed=20261017)
    e2.set_moments(emom, mmom)
    e2.commit()
    for mode in ('M', 'H'):
        e1.mc_sweeps(mode, 25, 600.0)
        e2.mc_sweeps(mode, 25, 600.0)
        a, _, _ = e1.get_moments()
        b, _, _ = e2.get_moments()
        assert np.array_equal(a, b), mode
    # and back to spin dynamics on the lattice layout
    e1.sd_steps(10)
    e2.sd_steps(10)
    assert np.abs(e1.get_moments()[0] - e2.get_moments()[0]).max() <= 1e-13


@pytest.mark.parametrize('mode', ['M', 'H'])
def test_cooperative_and_persistent_sweeps_run_the_same_chain(mode, monkeypatch):
    """FeCo B2 (z = 258, many small colour classes): the sub-warp cooperative update (warp-shuffle reduction of the
    Heisenberg sum) and the single cooperative launch with grid barriers between colours make the same draws and the
    same decisions as the one-thread-per-update colour launches; only the summation order of the field differs."""
    from uppasd_b200 import host
    monkeypatch.setenv('ASD_RESIDENT', '0')      # the launch-per-colour family is under test here, not the small-system kernel
    inp, S = _system('feco', 3)
    rng = np.random.default_rng(17)
    e0 = rng.normal(size=(3, S['Natom'], 3)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    out = {}
    for name, env in (('thread', dict(ASD_MC_PERSISTENT='0', ASD_MC_LPA='1')), ('subwarp', dict(ASD_MC_PERSISTENT='0', ASD_MC_LPA='16')),
                      ('persistent', dict())):
        for k in ('ASD_MC_PERSISTENT', 'ASD_MC_LPA'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        e = host.engine_from_system(S, orc.CONST, temp=900.0, seed=33)
        e.mc_sweeps(mode, 0, 900.0)             # builds the colour-major layout, runs nothing
        n0 = e.launch_count()
        e.mc_sweeps(mode, 4, 900.0)
        e.mc_sweeps(mode, 3, 900.0, first_sweep=5)
        nl = e.launch_count() - n0
        out[name] = (e.get_moments()[0], nl)
        assert np.abs(np.linalg.norm(out[name][0], axis=0) - 1.0).max() < 1e-12
    lay, ncol, _ = e.mc_colouring()
    assert lay == 0 and ncol >= 12
    assert out['persistent'][1] == 2 and out['thread'][1] == 7 * ncol          # one launch per call vs one per colour
    assert np.abs(out['thread'][0] - e0).max() > 0.1                            # the chain moved
    for name in ('subwarp', 'persistent'):
        assert np.abs(out[name][0] - out['thread'][0]).max() <= 1e-9, name


@pytest.mark.parametrize('name', ['bccfe', 'kagome', 'feco'])
@pytest.mark.parametrize('mode', ['M', 'H'])
def test_resident_monte_carlo_runs_the_same_chain(name, mode, monkeypatch):
    """Small systems: all colours of all sweeps of a call in ONE launch with the ensemble's state in shared memory
    (mc_resident_kernel, the default at this size) against one launch per colour with one thread per update: the same
    draws, the same field summation order, hence the same chain -- to the last bit or two: mc_update_site is inlined into both
    kernels and the compiler contracts multiply-adds per copy (observed: 2.2e-16 on 7 % of the components); one launch per call."""
    from uppasd_b200 import host
    inp, S = _system(name, 3)
    rng = np.random.default_rng(23)
    e0 = rng.normal(size=(3, S['Natom'], 3)); e0 /= np.sqrt((e0 ** 2).sum(axis=0))
    S['emom'] = np.asfortranarray(e0)
    S['emomM'] = np.asfortranarray(e0 * S['mmom'][None])
    T = 400.0
    out = {}
    for tag, env in (('resident', dict()), ('launches', dict(ASD_RESIDENT='0', ASD_MC_PERSISTENT='0', ASD_MC_LPA='1'))):
        for k in ('ASD_RESIDENT', 'ASD_MC_PERSISTENT', 'ASD_MC_LPA'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        e = host.engine_from_system(S, orc.CONST, temp=T, seed=9)
        e.mc_sweeps(mode, 0, T)
        n0 = e.launch_count()
        e.mc_sweeps(mode, 5, T)
        e.mc_sweeps(mode, 4, T, first_sweep=6)
        out[tag] = (e.get_moments()[0], e.launch_count() - n0)
    _, ncol, _ = e.mc_colouring()
    # (each count includes the one conversion launch of get_moments)
    assert out['resident'][1] == 2 + 1 and out['launches'][1] == 9 * ncol + 1, (out['resident'][1], out['launches'][1], ncol)
    assert np.abs(out['resident'][0] - out['launches'][0]).max() <= 1e-14
    assert np.abs(out['resident'][0] - e0).max() > 0.1


def test_bccfe_temperature_scan_magnetisation_binder_and_tc():
    """BASELINE config 2 in miniature (Tc scan, ensembles in parallel): heat-bath sweeps of tests/bccFe (6^3 cells) at six
    temperatures across the transition, 16 ensembles on the GPU (colour-parallel, resident kernel) against the oracle's
    random-sequential chain (4 ensembles).  <|M|> and the Binder cumulant U4 = 1 - <m^4> / (3 <m^2>^2) agree within the
    statistical error bars at every temperature, and the temperature at which <|M|> has dropped to half the moment (a
    finite-size Tc estimate) agrees within 40 K."""
    from uppasd_b200 import host
    temps = [700.0, 800.0, 900.0, 1000.0, 1100.0, 1300.0]
    inp, S = _system('bccfe', 16)
    inp4, S4 = _system('bccfe', 4)
    m0 = float(S['mmom'][0, 0])
    e = host.engine_from_system(S, orc.CONST, temp=temps[0], seed=41)
    gpu_m, ref_m = [], []
    sweep = 1
    for T in temps:
        e.mc_sweeps('H', 400, T, first_sweep=sweep)
        sweep += 400
        gm = []
        for r in range(300):
            e.mc_sweeps('H', 5, T, first_sweep=sweep)
            sweep += 5
            gm.append(np.sqrt(((e.measure() / S['Natom']) ** 2).sum(axis=0)))
        gm = np.array(gm)                                   # (samples, 16)
        rm, _, _ = orc.mc_run(S4, 'H', T, 1200, seed=7, sample_every=5, burn=300)      # (samples, 4)

        def stats(x):
            m1, m2, m4 = x.mean(axis=0), (x ** 2).mean(axis=0), (x ** 4).mean(axis=0)
            return m1, 1.0 - m4 / (3.0 * m2 ** 2)

        (g1, gu), (r1, ru) = stats(gm), stats(rm)
        sig_m = np.sqrt(g1.var() / len(g1) + r1.var() / len(r1))
        sig_u = np.sqrt(gu.var() / len(gu) + ru.var() / len(ru))
        assert abs(g1.mean() - r1.mean()) < 5 * sig_m + 0.03, (T, g1.mean(), r1.mean(), sig_m)
        assert abs(gu.mean() - ru.mean()) < 5 * sig_u + 0.02, (T, gu.mean(), ru.mean(), sig_u)
        gpu_m.append(g1.mean())
        ref_m.append(r1.mean())

    def t_half(ms):
        ms = np.array(ms) / m0
        for a in range(len(temps) - 1):
            if ms[a] >= 0.5 > ms[a + 1]:
                return temps[a] + (ms[a] - 0.5) / (ms[a] - ms[a + 1]) * (temps[a + 1] - temps[a])
        raise AssertionError(('no crossing', list(ms)))

    assert gpu_m[0] > 0.6 * m0 and gpu_m[-1] < 0.35 * m0, gpu_m
    assert abs(t_half(gpu_m) - t_half(ref_m)) < 40.0, (t_half(gpu_m), t_half(ref_m), gpu_m, ref_m)
