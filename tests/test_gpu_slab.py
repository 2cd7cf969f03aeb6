"""Slab decomposition + fused halo push: a supercell cut into z-slabs must reproduce the undecomposed run bit for
bit (same kernels, same summation order, noise keyed by the global atom index) -- T = 0 and thermal, both solvers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bcc(ncell, solver, temp, nslabs=0, mens=1, damping=0.5, seed=77, spread=False):
    """list of engines: one undecomposed (nslabs = 0) or `nslabs` slabs living in this process, on device 0 or
    (spread) round-robin over all visible devices -> peer stores cross NVLink"""
    from uppasd_b200 import capi
    ndev = capi.load().asd_device_count() if spread else 1
    import bench
    from uppasd_b200 import host, lattice, slab
    B, CONST = bench.BCC, bench.CONST
    ns, ca, cs, sh = lattice.stencil(B['cell'], B['bas'], B['atype'], np.array([4]), B['shells'][None], 1, np.ones((1, 4), dtype=int))
    cp = lattice.couplings(ns, ca, sh, B['atype'], B['J'][None, None, :], B['mom'], CONST['mry'], CONST['mub'])
    H = slab.halo_depth(cs, ns)
    assert H == 2
    out = []
    for g in range(max(nslabs, 1)):
        e = host.Engine(g % ndev)
        e.set_constants(CONST['gama'], CONST['k_bolt'], CONST['mub'], CONST['mry'])
        nz = ncell[2] // max(nslabs, 1)
        n = 2 * ncell[0] * ncell[1] * nz
        e.set_system(n, mens, 2, (np.arange(n, dtype=np.int32) % 2) + 1)
        if nslabs:
            e.set_slab(nslabs, g, H)
        e.build_lattice_table(0, 2, ncell, ('P', 'P', 'P'), ns, ca, cs, cp)
        e.set_llg(solver, 1e-16, landeg=1.0, lambda1=damping, temp=temp, seed=seed)
        e.commit()
        out.append(e)
    if nslabs:
        for g, e in enumerate(out):
            e.slab_connect_local(out[(g - 1) % nslabs], out[(g + 1) % nslabs])
    for e in out:
        e.init_moments_tilted(0.3, B['mom'])
    return out


def _gather(engines):
    return np.concatenate([e.get_moments()[0] for e in engines], axis=1)


@pytest.mark.parametrize('solver', [1, 5])
@pytest.mark.parametrize('temp', [0.0, 300.0])
@pytest.mark.parametrize('nslabs', [1, 2, 4])
@pytest.mark.parametrize('ncell', [(32, 6, 16), (64, 6, 16)])
def test_slab_matches_undecomposed(solver, temp, nslabs, ncell):
    """(32, 6, 16): one periodic 32-cell x-run -> staged kernel; (64, 6, 16): run-compressed register-blocked kernel on
    1024-slot super-brick tiles (boundary / interior split in those tiles)"""
    ref = _bcc(ncell, solver, temp)[0]
    sl = _bcc(ncell, solver, temp, nslabs=nslabs)
    assert ref.layout_info()['runs'] == (4 if ncell[0] == 64 else 0)
    assert all(e.layout_info()['runs'] == ref.layout_info()['runs'] for e in sl)
    assert np.array_equal(_gather(sl), ref.get_moments()[0])          # same tilted start (global index)
    done = 0
    for n in (1, 3, 20):
        ref.sd_steps(n, first_step=done + 1)
        # interleave the slabs step by step: inside one process a slab's halo wait needs its neighbours' launches queued
        for s in range(n):
            for e in sl:
                e.sd_steps(1, first_step=done + 1 + s)
        done += n
        a, b = _gather(sl), ref.get_moments()[0]
        assert np.array_equal(a, b), (solver, temp, nslabs, done, np.abs(a - b).max())
    for e in sl:
        ep, err = e.slab_status()
        # one exchange per stage + the initial push (+ one refresh of the halo spins per asd_sd_steps call on the moment planes)
        assert err == 0 and ep == 1 + 2 * done + (done if e.layout_info()['planes'] else 0)


@pytest.mark.parametrize('solver', [1, 5])
@pytest.mark.parametrize('nslabs', [2, 4])
def test_slab_on_moment_planes(solver, nslabs):
    """Slabs whose layout takes the moment-plane instantiations of the run kernel (asd_runs.cuh): a boundary atom's emomM is pushed
    into the neighbours' planes next to its spin, and the planes of the halo are derived after the neighbours' first push has
    landed.  Thermal run, bit for bit against the undecomposed supercell; a Monte Carlo interlude invalidates and rebuilds the planes."""
    ncell = (64, 8, 32)
    ref = _bcc(ncell, solver, 300.0)[0]
    sl = _bcc(ncell, solver, 300.0, nslabs=nslabs)
    assert ref.layout_info()['planes'] == 1 and all(e.layout_info()['planes'] == 1 for e in sl), [e.layout_info() for e in sl]
    done = 0
    for n in (1, 2, 9):
        ref.sd_steps(n, first_step=done + 1)
        for s in range(n):
            for e in sl:
                e.sd_steps(1, first_step=done + 1 + s)
        done += n
        a, b = _gather(sl), ref.get_moments()[0]
        assert np.array_equal(a, b), (solver, nslabs, done, np.abs(a - b).max())
    for e in sl:
        assert e.slab_status()[1] == 0
    # the stages exchanged emomM only; the SPINS of the halos were refreshed when asd_sd_steps returned: field evaluation reads them
    b_ref = ref.effective_field(energy=False)[0]
    b_sl = np.concatenate([e.effective_field(energy=False)[0] for e in sl], axis=1)
    assert np.abs(b_sl - b_ref).max() <= 1e-13 * np.abs(b_ref).max()


def test_slab_observables_and_ensembles():
    ncell = (32, 4, 8)
    ref = _bcc(ncell, 1, 300.0, mens=3)[0]
    sl = _bcc(ncell, 1, 300.0, nslabs=2, mens=3)
    for s in range(10):
        ref.sd_steps(1, first_step=s + 1)
        for e in sl:
            e.sd_steps(1, first_step=s + 1)
    m_ref = ref.measure()
    m_sl = sum(e.measure() for e in sl)
    assert np.allclose(m_sl, m_ref, rtol=1e-13, atol=1e-9)
    assert np.array_equal(_gather(sl), ref.get_moments()[0])


def test_slab_errors_are_loud():
    from uppasd_b200 import host
    sl = _bcc((32, 4, 8), 1, 0.0, nslabs=2)
    with pytest.raises(host.AsdError):
        sl[0].set_mc_layout(0)                  # a slab cannot re-order its atoms colour by colour
    e = host.Engine(0)
    e.set_system(2 * 32 * 4 * 3, 1, 2, (np.arange(2 * 32 * 4 * 3, dtype=np.int32) % 2) + 1)
    e.set_slab(2, 0, 2)
    import bench
    from uppasd_b200 import lattice
    B, CONST = bench.BCC, bench.CONST
    ns, ca, cs, sh = lattice.stencil(B['cell'], B['bas'], B['atype'], np.array([4]), B['shells'][None], 1, np.ones((1, 4), dtype=int))
    cp = lattice.couplings(ns, ca, sh, B['atype'], B['J'][None, None, :], B['mom'], CONST['mry'], CONST['mub'])
    with pytest.raises(host.AsdError):
        e.build_lattice_table(0, 2, (32, 4, 7), ('P', 'P', 'P'), ns, ca, cs, cp)   # 7 planes do not split in two


def test_slab_across_devices():
    """the same check with the slabs on different GPUs of the box (skipped on a single-GPU box)"""
    from uppasd_b200 import capi
    ndev = capi.load().asd_device_count()
    if ndev < 2:
        pytest.skip('needs at least two GPUs')
    ncell = (64, 8, 4 * ndev)
    ref = _bcc(ncell, 1, 300.0)[0]
    sl = _bcc(ncell, 1, 300.0, nslabs=ndev, spread=True)
    for s in range(25):
        ref.sd_steps(1, first_step=s + 1)
        for e in sl:
            e.sd_steps(1, first_step=s + 1)
    assert np.array_equal(_gather(sl), ref.get_moments()[0])
    assert all(e.slab_status()[1] == 0 for e in sl)


def _proper(e, col):
    """no atom shares its colour with any atom of its neighbour list"""
    nlist, nsize, _ = e.get_table(0)
    for i in range(e.N):
        nb = nlist[:nsize[i % len(nsize)] if len(nsize) < e.N else nsize[i], i] - 1
        nb = nb[nb != i]
        if (col[nb] == col[i]).any():
            return False
    return True


@pytest.mark.parametrize('ncell', [(32, 6, 16), (12, 5, 8), (33, 4, 6)])
def test_periodic_colouring_is_proper(ncell):
    e = _bcc(ncell, 1, 0.0)[0]
    e.set_mc_layout(1)
    lay, ncol, per = e.mc_colouring()
    assert lay == 1 and 2 <= ncol < 64
    for a in range(3):
        # periodic directions: the period divides the extent (periods down to 2 are admissible when no stencil shift aliases an
        # atom with itself in the quotient: bcc Fe with 4 shells takes the 8-colouring of period (2, 2, 2) where it divides)
        assert ncell[a] % per[a] == 0 and per[a] >= 2
    col = e.get_mc_colours()
    assert col.min() == 0 and col.max() == ncol - 1
    assert _proper(e, col)
    # the colour-major layout colours the same graph (greedy on the actual lists)
    e2 = _bcc(ncell, 1, 0.0)[0]
    assert _proper(e2, e2.get_mc_colours())


@pytest.mark.parametrize('mode', ['M', 'H'])
@pytest.mark.parametrize('nslabs', [1, 2, 4])
def test_slab_mc_matches_undecomposed(mode, nslabs):
    """Monte Carlo on slabs: one fused halo exchange per colour; the chain is the one of the undecomposed lattice layout
    bit for bit (global colouring, draws keyed by the global atom index)."""
    ncell = (32, 6, 16)
    ref = _bcc(ncell, 1, 0.0, mens=2)[0]
    ref.set_mc_layout(1)
    sl = _bcc(ncell, 1, 0.0, nslabs=nslabs, mens=2)
    lay, ncol, per = ref.mc_colouring()
    for e in sl:
        assert e.mc_colouring() == (1, ncol, per)
    gcol = ref.get_mc_colours()
    assert np.array_equal(np.concatenate([e.get_mc_colours() for e in sl]), gcol)
    done = 0
    for n in (1, 2, 5):
        ref.mc_sweeps(mode, n, 600.0, first_sweep=done + 1)
        for s in range(n):
            for e in sl:
                e.mc_sweeps(mode, 1, 600.0, first_sweep=done + 1 + s)
        done += n
        a, b = _gather(sl), ref.get_moments()[0]
        assert np.array_equal(a, b), (mode, nslabs, done, np.abs(a - b).max())
    assert np.abs(np.linalg.norm(ref.get_moments()[0], axis=0) - 1.0).max() < 1e-12
    for e in sl:
        ep, err = e.slab_status()
        assert err == 0 and ep == 1 + ncol * done
    # SD steps continue on the same buffers after the sweeps (halos of `cur` are current)
    ref.sd_steps(3, first_step=1)
    for s in range(3):
        for e in sl:
            e.sd_steps(1, first_step=1 + s)
    assert np.array_equal(_gather(sl), ref.get_moments()[0])


def test_lattice_mc_layout_agrees_with_colour_major_on_observables():
    """two different proper colourings = two Markov chains with the same stationary distribution: <|m|> at 600 K agrees
    within the statistical error (bcc Fe, T well below Tc)."""
    ncell = (32, 8, 8)
    res = []
    for layout in (0, 1):
        e = _bcc(ncell, 1, 0.0, mens=4, seed=5 + layout)[0]
        e.set_mc_layout(layout)
        e.mc_sweeps('H', 300, 600.0)
        ms = []
        for b in range(40):
            e.mc_sweeps('H', 10, 600.0, first_sweep=301 + 10 * b)
            ms.append(np.linalg.norm(e.measure(), axis=0) / e.N)
        ms = np.array(ms)                      # (40, M)
        res.append((ms.mean(), ms.mean(axis=0).std() / np.sqrt(ms.shape[1]) + ms.std() / np.sqrt(ms.size)))
    (m0, s0), (m1, s1) = res
    assert abs(m0 - m1) < 5 * (s0 + s1) + 2e-3, res


@pytest.mark.parametrize('solver', [1, 5])
def test_single_slab_whose_ring_neighbour_is_itself(solver):
    """nslabs = 1: the slab machinery (halo wait, boundary / interior launches, side-stream exchange of the moment planes, refresh of
    the halo spins when asd_sd_steps returns) on one periodic supercell whose lower and upper neighbour are the slab itself -- bit
    for bit the undecomposed run, fields included."""
    import bench
    ncell = (64, 8, 16)
    ref, n = bench.bcc_engine(ncell, solver, 300.0, 0.5, 2, 0, 0)
    sl, _ = bench.bcc_engine(ncell, solver, 300.0, 0.5, 2, 0, 0, slab=(1, 0, None))
    assert sl.layout_info()['planes'] == 1 and ref.layout_info()['planes'] == 1
    for nst in (1, 4, 9):
        ref.sd_steps(nst, first_step=1)
        sl.sd_steps(nst, first_step=1)
        assert np.array_equal(ref.get_moments()[0], sl.get_moments()[0]), (solver, nst)
    assert sl.slab_status()[1] == 0
    b_ref = ref.effective_field(energy=False)[0]
    b_sl = sl.effective_field(energy=False)[0]
    assert np.abs(b_sl - b_ref).max() <= 1e-13 * np.abs(b_ref).max()
    ref.close(); sl.close()
