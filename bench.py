#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native UppASD hot path.

Metric (BASELINE.json): atom-steps/s of the stochastic LLG step (semi-implicit midpoint, SDEalgh 1) on the
bcc Fe 128x128x128 supercell (4 194 304 spins, 4 exchange shells, z = 50, reduced Hamiltonian), T = 300 K.
One "step" = one full LLG time step of every spin (field / predictor / field / corrector / moment update).

  python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun, one rank/GPU)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: restated CPU path on the host cores

N > 1 shards independent ensembles (Mensemble = N, one per GPU, no data-path communication): weak scaling.
All timing is on the device (CUDA events on the engine's stream), max over ranks, after a barrier.

The one JSON line also carries `secondary`: the Monte Carlo sweep rate of the same supercell (Metropolis and heat bath,
attempts/s against the 256 B/attempt figure of SURVEY 8d), BASELINE configs 3 (FeCo random alloy, Mensemble 8: Monte Carlo
and LLG) and 4 (2-D triangular Heisenberg + DMI lattice: LLG and heat bath) through the run-directory driver, and, for N > 1,
the slab-decomposed single supercell with the fused NVLink halo push (BASELINE config 5: bcc 512 x 512 x 256 on 8 GPUs,
256^3 below), strong scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONST = dict(gama=1.760859644e11, k_bolt=1.38064852e-23, mub=9.274009994e-24, mry=2.179872325e-21)
# tests/bccFe/{posfile,momfile,jfile}: bcc Fe, 4 shells (mRy), m = 2.23 mu_B
BCC = dict(cell=np.eye(3), bas=np.array([[0.0, 0.5], [0.0, 0.5], [0.0, 0.5]]), atype=np.array([1, 1]),
           mom=np.array([2.23, 2.23]),
           shells=np.array([[0.5, 0.5, 0.5], [1.0, 0.0, 0.0], [1.0, 1.0, 0.0], [1.5, 0.5, 0.5]]),
           J=np.array([1.33767484769984, 0.75703576545650, -0.05975437643846, -0.08819834160658]))
B_ALG = {1: (256.0, 280.0), 5: (280.0, 304.0)}  # algorithmic bytes/atom of stage 1, stage 2 (z = 50)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU while the timed region runs (NVML, every ~5 ms; nvidia-smi as
    a fallback)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self.stop_flag = index, [], [], set(), False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and vis.split(',')[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))   # before the timed region
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is not None:
            nv = self.nv
            bits = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown,
                    'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
            # the first sample is taken the moment the thread starts (the timed region of a 20-step run is ~10 ms: with eight ranks
            # querying NVML at once a sample takes several ms) and the loop always completes the sample it began
            while True:
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, b in bits.items():
                        if r & b:
                            self.reasons.add(k)
                except Exception:
                    pass
                if self.stop_flag:
                    break
                time.sleep(0.002)
            return
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while True:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                r = [x.strip() for x in out.strip().split(',')]
                if len(r) >= 6:
                    self.sm.append(float(r[0])); self.mx.append(float(r[1]))
                    for i in range(4):
                        if r[2 + i].lower().startswith('active'):
                            self.reasons.add(names[i])
            except Exception:
                pass
            if self.stop_flag:
                break
            time.sleep(0.05)

    def summary(self):
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': max(self.mx) if self.mx else None,
                'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': 'nvml' if self.nv is not None else 'nvidia-smi'}


def bcc_engine(ncell, solver, temp, damping, mensemble, ens_offset, device, slab=None, reduced=True):
    """bcc Fe supercell built entirely on the device (tables + tilted start), through the C ABI.
    slab = (world, rank, dist): this engine holds one z-slab of the supercell and is connected to its ring neighbours."""
    from uppasd_b200 import host, lattice
    from uppasd_b200 import slab as slabmod
    nn = np.array([4])
    red = BCC['shells'][None]                       # (NT=1, 4, 3)
    ns, ca, cs, sh = lattice.stencil(BCC['cell'], BCC['bas'], BCC['atype'], nn, red, 1, np.ones((1, 4), dtype=int))
    cp = lattice.couplings(ns, ca, sh, BCC['atype'], BCC['J'][None, None, :], BCC['mom'], CONST['mry'], CONST['mub'])
    e = host.Engine(device)
    e.set_constants(CONST['gama'], CONST['k_bolt'], CONST['mub'], CONST['mry'])
    nz = ncell[2]
    if slab is not None:
        world, rank, dist = slab
        halo = slabmod.halo_depth(cs, ns)
        _, nz = slabmod.slab_planes(ncell[2], world, rank, halo)
    n = 2 * ncell[0] * ncell[1] * nz
    aham = (np.arange(n, dtype=np.int32) % 2) + 1
    if reduced:
        e.set_system(n, mensemble, 2, aham)
    else:
        e.set_system(n, mensemble, n, None)          # do_reduced N: one coupling row per atom (ncoup(z, Natom))
    if slab is not None:
        e.set_slab(world, rank, halo)
    e.build_lattice_table(0, 2, ncell, ('P', 'P', 'P'), ns, ca, cs, cp)
    e.set_llg(solver, 1e-16, landeg=1.0, lambda1=damping, temp=temp, seed=20261017)
    e.set_ensemble_offset(ens_offset)
    e.commit()
    if slab is not None:
        slabmod.connect_ring(e, world, rank, dist)
    e.init_moments_tilted(0.1, BCC['mom'])
    return e, n


def oracle_bcc(ncell, mensemble=1):
    """Same system through the CPU oracle (test infrastructure; used only for the CPU baseline legs)."""
    from oracle import orc
    inp = dict(ncell=tuple(ncell), bc=('P', 'P', 'P'), cell=BCC['cell'], sym=1, mensemble=mensemble, do_reduced='Y',
               do_sortcoup='N', map_multiple=False, hfield=(0.0, 0.0, 0.0))
    ex = (np.array([4], dtype=np.int32), BCC['shells'][None].copy(), BCC['J'][None, None, :].copy(),
          np.ones((1, 4), dtype=np.int32))
    aemom = np.array([[1.0, 1.0], [0.0, 0.0], [0.0, 0.0]])
    S = orc.build_system(inp, BCC['bas'], BCC['atype'], BCC['mom'], aemom, np.array([2.0, 2.0]), ex)
    n = S['Natom']
    i = np.arange(1, n + 1, dtype=np.float64)
    h = np.modf(i * 0.6180339887)[0]
    e = np.stack([np.ones(n), 0.1 * np.sin(2 * np.pi * h), 0.1 * np.cos(2 * np.pi * h)])
    e /= np.sqrt((e ** 2).sum(axis=0))
    for k in range(mensemble):
        S['emom'][:, :, k] = e
        S['emomM'][:, :, k] = e * S['mmom'][:, k]
    return S


def cpu_leg(ncell, solver, temp, damping, steps, warmup, settle_s=0.0, one_thread_steps=0):
    """Times the restated reference CPU path (oracle, OpenMP over all host cores) on a bounded sample; optionally the same
    system on ONE thread as well (SURVEY 8d asks for both).  Returns (rate, ms/step, atoms, threads, one-thread rate | None)."""
    from oracle import orc
    # all host cores, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1 for its workers)
    threads = orc.set_num_threads(os.cpu_count() or 1)
    S = oracle_bcc(ncell)
    n = S['Natom']
    st = orc.SdState(S, solver, 1e-16, damping, temp=temp)
    rng = np.random.default_rng(1)
    g = np.asfortranarray(rng.normal(size=(3, n, 1))) if temp > 0 else None   # noise generation not timed (favours the CPU)
    tw = time.perf_counter()
    w = 0
    # settle_s: under a multi-rank launcher the other ranks are still starting their interpreters (and exit at once); keep
    # warming up until they are gone so that the timed steps have the host cores to themselves
    while w < warmup or time.perf_counter() - tw < settle_s:
        st.step(gauss=g)
        w += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step(gauss=g)
    dt = time.perf_counter() - t0
    one = None
    if one_thread_steps > 0:
        orc.set_num_threads(1)
        t1 = time.perf_counter()
        for _ in range(one_thread_steps):
            st.step(gauss=g)
        one = n * one_thread_steps / (time.perf_counter() - t1)
        orc.set_num_threads(threads)
    return n * steps / dt, dt / steps * 1e3, n, threads, one


def bind_near_gpu(torch, local):
    """Multi-rank runs: pin this process to the CPUs NVML reports as local to its GPU (intersected with the CPUs the container
    allows), so that the pinned host buffers of the end-to-end leg are allocated on the GPU's own NUMA node and the state copies
    of the eight ranks do not all cross one socket link.  Returns a short description for the JSON line, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {64 * w + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = near & allowed
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            return {'cpus_near_gpu': len(near), 'bound_to': len(target), 'allowed': len(allowed)}
        return {'cpus_near_gpu': len(near), 'bound_to': 0, 'allowed': len(allowed)}
    except Exception as ex:   # no NVML, no permission: run unbound
        return {'error': repr(ex)[:120]}


def traffic_from_profiles(kernel_key):
    """dram bytes per launch of the dominant kernel from the committed ncu summary (profiles/traffic.json), or None"""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel_key)
        except Exception:
            return None
    return None


def _config3_dir(d, ncell, mens):
    """BASELINE config 3 as the reference's example states it (examples/Mappings/FeCo/random: bcc primitive cell, 50/50 random
    occupancy, z = 258), materialised from tests/golden/feco_random.json with the supercell / ensemble count of the run"""
    fx = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'feco_random.json')))
    for k, v in fx['raw'].items():
        open(os.path.join(d, k), 'w').write(v)
    drop = ('ncell', 'mensemble', 'ip_mode', 'mode', 'sdealgh', 'temp')
    lines = [l for l in fx['raw']['inpsd.dat'].splitlines() if not (l.split() and l.split()[0].lower() in drop)]
    lines += ['ncell %d %d %d' % tuple(ncell), 'mensemble %d' % mens, 'ip_mode N', 'mode M', 'sdealgh 1', 'temp 600']
    open(os.path.join(d, 'inpsd.dat'), 'w').write('\n'.join(lines) + '\n')
    return os.path.join(d, 'inpsd.dat')


def _config4_dir(d, ncell, mens):
    """BASELINE config 4: 2-D triangular lattice, nearest-neighbour Heisenberg exchange + interfacial DMI + uniaxial anisotropy +
    field along -z (the ingredients of examples/SpecialFeatures/SkyrmionLattice), reduced Hamiltonian"""
    a1, a2 = (1.0, 0.0, 0.0), (-0.5, 0.8660254037844386, 0.0)
    nbr = [(1, 0), (0, 1), (-1, -1), (-1, 0), (0, -1), (1, 1)]
    vec = lambda n: np.array([n[0] * a1[0] + n[1] * a2[0], n[0] * a1[1] + n[1] * a2[1], 0.0])
    open(os.path.join(d, 'posfile'), 'w').write('1 1 0.0 0.0 0.0\n')
    open(os.path.join(d, 'momfile'), 'w').write('1 1 1.5 0.1 0.05 1.0\n')
    with open(os.path.join(d, 'jfile'), 'w') as fh:
        for n in nbr:
            fh.write('1 1 %.10f %.10f %.10f 1.0\n' % tuple(vec(n)))
    with open(os.path.join(d, 'dmfile'), 'w') as fh:
        for n in nbr:
            r = vec(n)
            dm = 0.35 * np.cross([0.0, 0.0, 1.0], r / np.linalg.norm(r))
            fh.write('1 1 %.10f %.10f %.10f %.10f %.10f %.10f\n' % (*r, *dm))
    open(os.path.join(d, 'kfile'), 'w').write('1 1 0.05 0.0 0.0 0.0 1.0 0.0\n')
    open(os.path.join(d, 'inpsd.dat'), 'w').write(
        'simid skyrm_2D\nncell %d %d 1\nBC P P 0\ncell %.10f %.10f %.10f\n     %.10f %.10f %.10f\n     0.0 0.0 1.0\nSym 0\n'
        'posfile ./posfile\nmomfile ./momfile\nexchange ./jfile\ndm ./dmfile\nanisotropy ./kfile\ndo_reduced Y\nMensemble %d\n'
        'Initmag 3\nSDEalgh 1\nmode S\ntemp 10\nhfield 0.0 0.0 -2.5\ndamping 0.3\ntimestep 1.0d-16\nNstep 10\ndo_avrg N\n'
        % (ncell[0], ncell[1], *a1, *a2, mens))
    return os.path.join(d, 'inpsd.dat')


def config_blocks(world, rank, local, dist, torch, peak):
    """secondary.config3 / secondary.config4: the two BASELINE configurations that are not bcc Fe, through the run-directory
    driver (uppasd_b200/driver.py), one independent copy per GPU (weak scaling): attempts/s and atom-steps/s summed over ranks"""
    import tempfile
    import warnings
    from uppasd_b200 import driver
    out = {}

    def agg(ms):
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        try:
            sim = driver.Simulation(_config3_dir(tempfile.mkdtemp(), (32, 32, 32), 8), device=local, seed=1000 + rank)
            e, n, m = sim.engine, sim.natom, sim.mens
            z = int(sim.tables['nlist'].shape[0])
            blk = {'workload': 'FeCo random alloy (bcc primitive cell, 50/50, z = %d, one coupling row per atom) 32^3 = %d atoms x %d '
                               'ensembles per GPU, T = 600 K' % (z, n, m)}
            for mode, key in (('M', 'metropolis'), ('H', 'heat_bath')):
                e.mc_sweeps(mode, 3, 600.0)
                ms = agg(e.time_mc_sweeps(mode, 20, 600.0))
                rate = world * n * m * 20 / (ms * 1e-3)
                blk[key] = {'value': rate, 'unit': 'attempts/s', 'ms_per_sweep': ms / 20,
                            'frac_of_peak': (56.0 + 12.0 * z) * rate / world / 1e9 / peak}
            e.sd_steps(3)
            ms = agg(e.time_sd_steps(20, first_step=4))
            rate = world * n * m * 20 / (ms * 1e-3)
            blk['llg'] = {'value': rate, 'unit': 'atom-steps/s', 'ms_per_step': ms / 20, 'frac_of_peak': (136.0 + 24.0 * z) * rate / world / 1e9 / peak}
            blk['note'] = ('frac_of_peak counts the per-atom tables (4z index + 8z coupling bytes) once per (atom, ensemble) and stage as SURVEY 8d '
                           'does; here they are shared by the %d ensembles and (%.0f MB) stay in the 126 MB L2, so a value above 1 is not HBM traffic'
                           % (m, 12.0 * z * n / 1e6))
            e.close()
            out['config3'] = blk
        except Exception as ex:
            out['config3'] = {'error': repr(ex)[:300]}
        try:
            sim = driver.Simulation(_config4_dir(tempfile.mkdtemp(), (1024, 1024), 2), device=local, seed=2000 + rank)
            e, n, m = sim.engine, sim.natom, sim.mens
            balg = 136.0 + 8 * 6 + 8 * 6 + 104
            blk = {'workload': '2-D triangular lattice 1024 x 1024 x %d ensembles per GPU, Heisenberg + interfacial DMI + uniaxial anisotropy '
                               '+ field, reduced Hamiltonian, T = 10 K' % m}
            sim.relax('S', nstep=10, temperature=10.0, timestep=1e-16, damping=0.3)
            ms = agg(e.time_sd_steps(100, first_step=1000))
            rate = world * n * m * 100 / (ms * 1e-3)
            blk['llg'] = {'value': rate, 'unit': 'atom-steps/s', 'ms_per_step': ms / 100, 'alg_bytes_per_atom_step': balg,
                          'frac_of_peak': balg * rate / world / 1e9 / peak}
            e.mc_sweeps('H', 3, 10.0)
            ms = agg(e.time_mc_sweeps('H', 20, 10.0))
            rate = world * n * m * 20 / (ms * 1e-3)
            blk['heat_bath'] = {'value': rate, 'unit': 'attempts/s', 'ms_per_sweep': ms / 20}
            e.close()
            out['config4'] = blk
        except Exception as ex:
            out['config4'] = {'error': repr(ex)[:300]}
    return out


def slab_block(a, world, rank, local, dist, torch, out, warmup, steps):
    """BASELINE config 5 next to the headline line: one supercell (512 x 512 x 256 on 8 GPUs, 256^3 below) cut into z-slabs.
    A watchdog prints the headline line without it if the ring does not come up (a hang must not cost the main number)."""
    import signal
    ncell = [512, 512, 256] if world >= 8 else [256, 256, 256]
    state = {'done': False}

    def bail():
        if state['done']:
            return
        if rank == 0 and out is not None:
            out.setdefault('secondary', {})['slab'] = {'error': 'timed out after 240 s'}
            print(json.dumps(out), flush=True)
        os._exit(0)
    timer = threading.Timer(240.0, bail)
    timer.daemon = True
    timer.start()
    blk = None
    try:
        e, n = bcc_engine(ncell, a.solver, a.temp, a.damping, 1, 0, local, slab=(world, rank, dist))
        sync = torch.zeros(1, device='cuda')

        def barrier():
            dist.all_reduce(sync)
            torch.cuda.synchronize()
            e.synchronize()
        e.sd_steps(warmup, first_step=1)
        barrier()
        ms = e.time_sd_steps(steps, first_step=warmup + 1)
        barrier()
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        nt = torch.tensor([float(n)], device='cuda', dtype=torch.float64)
        dist.all_reduce(nt)
        ntot = float(nt.item())
        _, err = e.slab_status()
        ef = torch.tensor([float(err)], device='cuda', dtype=torch.float64)
        dist.all_reduce(ef, op=dist.ReduceOp.MAX)
        peak, _ = peaks()
        b1, b2 = B_ALG[a.solver]
        rate = ntot * steps / (ms * 1e-3)
        blk = {'workload': 'bccFe %dx%dx%d (%d spins), ONE supercell in %d z-slabs, LLG solver %d, T=%g K' % (*ncell, int(ntot), world, a.solver, a.temp),
               'value': rate, 'unit': 'atom-steps/s', 'scaling': 'strong', 'ms_per_step': ms / steps, 'steps': steps, 'warmup': warmup,
               'halo': {'planes': 2, 'bytes_per_exchange_per_side': 2 * ncell[0] * ncell[1] * 2 * (24 if e.layout_info().get('planes') else 32),
                        'exchanges_per_step': 2, 'payload': 'emomM (moment planes)' if e.layout_info().get('planes') else 'spins',
                        'transport': ('peer stores over NVLink (CUDA IPC) from a side-stream launch concurrent with the interior tiles, epoch flags'
                                      if e.layout_info().get('planes') else 'peer stores over NVLink from the boundary-tile launches (CUDA IPC), epoch flags'),
                        'timeout_flag': int(ef.item())},
               'step_frac_of_peak': (b1 + b2) * rate / 1e9 / (peak * world)}
        e.close()
    except Exception as ex:
        blk = {'error': repr(ex)[:300]}
    state['done'] = True
    timer.cancel()
    if rank == 0 and out is not None:
        out.setdefault('secondary', {})['slab'] = blk
        print(json.dumps(out), flush=True)
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--ncell', type=int, nargs=3, default=[128, 128, 128])
    ap.add_argument('--solver', type=int, default=1, choices=[1, 5])
    ap.add_argument('--full-ham', action='store_true', help='do_reduced N: per-atom coupling rows (SURVEY 8d: +16 z bytes per atom-step)')
    ap.add_argument('--temp', type=float, default=300.0)
    ap.add_argument('--damping', type=float, default=0.5)
    ap.add_argument('--decomp', default='ensemble', choices=['ensemble', 'slab'],
                    help='N > 1: one ensemble of the supercell per GPU (weak scaling, no communication) or one z-slab '
                         'of a single supercell per GPU with the fused NVLink halo push (strong scaling)')
    ap.add_argument('--cpu-ncell', type=int, nargs=3, default=None, help='CPU legs on a smaller supercell (default: the same one)')
    ap.add_argument('--no-secondary', action='store_true', help='skip the Monte Carlo and slab blocks')
    ap.add_argument('--no-cpu', action='store_true')
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    steps, warmup = a.steps, max(a.warmup, 3)
    cores = os.cpu_count() or 1
    workload = 'bccFe %dx%dx%d LLG %s (SDEalgh %d), T=%g K, damping %g, dt 1e-16, z=50, do_reduced %s' % (
        a.ncell[0], a.ncell[1], a.ncell[2], 'midpoint' if a.solver == 1 else 'Depondt', a.solver, a.temp, a.damping,
        'N' if a.full_ham else 'Y')

    cpu_ncell = a.cpu_ncell or a.ncell
    if a.impl == 'reference':
        if rank != 0:
            return
        os.environ['OMP_NUM_THREADS'] = str(cores)
        # the configuration it names (default bcc 128^3): every step is a full step of that supercell; the step count is
        # bounded (<= 20) so that the run ends within a few minutes on any host
        k = max(1, min(steps, 20))
        w = min(warmup, 5)       # a CPU step of the named supercell takes ~0.1 s: the requested warm-up (3 in the driver's run) is honoured
        v, ms, n, cores, one = cpu_leg(cpu_ncell, a.solver, a.temp, a.damping, k, w, settle_s=6.0 if world > 1 else 0.0, one_thread_steps=1)
        if a.cpu_ncell:
            workload += ' [CPU sample: bcc %dx%dx%d]' % tuple(cpu_ncell)
        sample = 'bcc %dx%dx%d (%d spins) x %d full steps of the same lattice / solver / temperature; restated Fortran loops ' \
                 '(oracle/), OpenMP static schedule over %d threads, noise array pre-generated' % (*cpu_ncell, n, k, cores)
        print(json.dumps({
            'impl': 'reference', 'metric': 'atom-steps/sec', 'value': v, 'unit': 'atom-steps/s', 'n_gpus': a.gpus,
            'steps': k, 'warmup': w, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': {'workload': workload},
            'cpu_baseline': {'value': v, 'unit': 'atom-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                             'one_thread': {'value': one, 'unit': 'atom-steps/s', 'steps': 1}},
            'e2e': {'value': v, 'unit': 'atom-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}))
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    numa = bind_near_gpu(torch, local) if world > 1 else None   # (N = 1 keeps every core for the CPU baseline leg)
    slab = (world, rank, dist) if a.decomp == 'slab' else None
    e, n = bcc_engine(a.ncell, a.solver, a.temp, a.damping, 1, 0 if slab else rank, local, slab=slab, reduced=not a.full_ham)
    sync = torch.zeros(1, device='cuda')

    def barrier():
        if world > 1:
            dist.all_reduce(sync)
        torch.cuda.synchronize()
        e.synchronize()

    e.sd_steps(warmup, first_step=1)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = e.launch_count()
    ms = e.time_sd_steps(steps, first_step=warmup + 1)
    launches = e.launch_count() - l0
    barrier()
    # per-kernel durations of the two stage kernels (CUDA events on the engine's stream), median of 8
    st1, st2 = [], []
    for r in range(8):
        _, (x, y) = e.time_sd_steps(0, first_step=warmup + steps + 1 + r, stages=True)
        st1.append(x); st2.append(y)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([ms], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * steps / (ms * 1e-3)

    # ---- end to end through the C ABI with HOST buffers (what a driver that owns the moments does, sd_mphase): H2D of the
    #      state from pinned host memory (asd_set_moments: emom + mmom), the measurement-phase loop asd_sd_run with a sample
    #      of sum M after EVERY step (reduced on the device into a sample ring, one D2H + one synchronisation at the end),
    #      D2H of the final unit vectors into pinned host memory (asd_get_moments(emom); |m| does not change with mompar 0
    #      and emomM = emom * mmom is the host's to form)
    emom, emomM, mmom = e.get_moments()

    def pinned(x):
        tt = torch.empty(x.size, dtype=torch.float64).pin_memory()
        v = tt.numpy().reshape(x.shape, order='F')
        v[...] = x
        return tt, v
    keep, h_e = pinned(emom)
    keep3, h_m = pinned(mmom)
    k2 = steps
    keep4, h_s = pinned(np.zeros((3, 1, k2), order='F'))
    del emomM
    # staging buffers of the three calls (device I/O buffers, sample ring, pinned landing zone) are allocated by a first, untimed pass
    e.set_moments(h_e, h_m)
    e.sd_run(2, first_step=9_000, sample_every=1)
    e.get_moments(out=(h_e, None, None))
    barrier()
    t0 = time.perf_counter()
    e.set_moments(h_e, h_m)
    samples = e.sd_run(k2, first_step=10_000, sample_every=1, out=h_s)
    e.get_moments(out=(h_e, None, None))
    e.synchronize()
    dt = time.perf_counter() - t0
    assert samples.shape[2] == k2 and np.isfinite(samples).all() and abs(np.sqrt((h_e[:, :8, 0] ** 2).sum(axis=0)) - 1.0).max() < 1e-9
    te = torch.tensor([dt], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = world * n * k2 / float(te.item())
    h2d = 32.0 * n / k2
    d2h = 24.0 * n / k2 + 24.0
    _, slab_err = e.slab_status()

    # ---- secondary: Monte Carlo on the same supercell (attempts/s, 256 B per attempt at z = 50) ----
    secondary = {}
    peak, how = peaks()
    if not a.no_secondary and not slab and not a.full_ham:
        try:
            mc = {}
            for mode, key in (('M', 'metropolis'), ('H', 'heat_bath')):
                e.mc_sweeps(mode, 5, a.temp)
                l0 = e.launch_count()
                msw = e.time_mc_sweeps(mode, 20, a.temp)
                tm = torch.tensor([msw], device='cuda', dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                rate = world * n * 20 / (float(tm.item()) * 1e-3)
                mc[key] = {'value': rate, 'unit': 'attempts/s', 'ms_per_sweep': float(tm.item()) / 20, 'sweeps': 20,
                           'launches_per_sweep': (e.launch_count() - l0) / 20.0,
                           'roofline': {'bound': 'hbm', 'alg_bytes_per_attempt': 256.0, 'achieved': 256.0 * rate / world / 1e9,
                                        'peak': peak, 'unit': 'GB/s', 'frac': 256.0 * rate / world / 1e9 / peak}}
            lay_mc, ncol, per = e.mc_colouring()
            mc['layout'] = {0: 'colour-major', 1: 'lattice tiles, one launch per colour', 2: 'block sweep'}[lay_mc]
            mc['colours'] = ncol
            mc['workload'] = 'bccFe %dx%dx%d, T=%g K, one ensemble per GPU' % (*a.ncell, a.temp)
            secondary['mc'] = mc
        except Exception as ex:                                   # the headline line must survive
            secondary['mc'] = {'error': repr(ex)[:300]}
        try:
            # the other integrator of the path (Depondt, SDEalgh 5) and the T = 0 rate of the headline solver, same supercell
            for key, alg, temp in (('llg_depondt', 5, a.temp), ('llg_t0', a.solver, 0.0)):
                e.set_llg(alg, 1e-16, landeg=1.0, lambda1=a.damping, temp=temp, seed=20261017)
                e.sd_steps(3, first_step=1)
                tm = torch.tensor([e.time_sd_steps(steps, first_step=4)], device='cuda', dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                rate = world * n * steps / (float(tm.item()) * 1e-3)
                bb = sum(B_ALG[alg])
                secondary[key] = {'value': rate, 'unit': 'atom-steps/s', 'ms_per_step': float(tm.item()) / steps, 'solver': alg, 'temp': temp,
                                  'step_alg_bytes_per_atom': bb, 'step_frac_of_peak': bb * rate / world / 1e9 / peak}
        except Exception as ex:
            secondary['llg_depondt'] = {'error': repr(ex)[:300]}
        secondary.update(config_blocks(world, rank, local, dist, torch, peak))

    if rank == 0:
        b1, b2 = B_ALG[a.solver]
        if a.full_ham:
            b1, b2 = b1 + 8.0 * 50, b2 + 8.0 * 50        # per-atom couplings add 8 z bytes per stage
        t2 = float(np.median(st2)) * 1e-3
        t1 = float(np.median(st1)) * 1e-3
        ach = b2 * n / t2 / 1e9
        lay = e.layout_info()
        if lay['runs']:
            kname = 'llg_runs_kernel<solver=%d,stage=2,tile=%d%s>' % (a.solver, lay['tile_slots'], ',planes' if lay.get('planes') else '')
            tables = '%.0f MB of gather lists + run-compressed tables' % ((4.0 * lay['ucap'] + 16.0 * (lay['union'] + 2) * lay['tile_slots'] / 128) / lay['tile_slots'] * n / 1e6 + 8.0 * n / 1e6)
        else:
            kname = 'llg_stage_kernel<solver=%d,stage=2,%s,%s>' % (a.solver, 'full' if a.full_ham else 'reduced', 'staged' if lay['staged'] else 'direct')
            tables = 'index tables %.0f MB' % ((7 * 16 + 24) * n / 1e6)
        par = ('ensemble-sharded x%d (one %d-spin ensemble per GPU, no communication)' % (world, n)) if not slab else \
              ('z-slabs x%d of one supercell, halo exchange by peer stores over NVLink (side-stream launch on moment-plane layouts, else fused '
               'into the boundary-tile launches)' % world)
        out = {
            'metric': 'atom-steps/sec', 'value': value, 'unit': 'atom-steps/s', 'n_gpus': world, 'steps': steps,
            'warmup': warmup, 'ms_per_step': ms / steps, 'higher_is_better': True,
            'scaling': 'strong' if slab else 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload, 'spins_per_gpu': n, 'ensembles': 1 if slab else world, 'parallelism': par,
                       'field_path': lay, 'numa': numa,
                       'l2_policy': 'inputs larger than L2 (%s + spins %s%.0f MB per GPU vs 126 MB L2)'
                                    % (tables, 'and moment planes ' if lay.get('planes') else '', (112.0 if lay.get('planes') else 64.0) * n / 1e6)},
            'clocks': sampler.summary(),
            'e2e': {'value': e2e, 'unit': 'atom-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': k2, 'note': 'asd_set_moments from pinned host memory (H2D of emom + mmom) + asd_sd_run(K, sample sum M '
                                         'after every step: device sample ring, one D2H + one sync) + asd_get_moments(emom) into '
                                         'pinned host memory; state copies amortised over the K steps'},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'kernel': kname,
                         'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                         'traffic': traffic_from_profiles(kname),
                         'peak_source': how, 'alg_bytes_per_atom': b2, 'kernel_ms': t2 * 1e3,
                         'stage1': {'alg_bytes_per_atom': b1, 'kernel_ms': t1 * 1e3, 'achieved': b1 * n / t1 / 1e9,
                                    'frac': b1 * n / t1 / 1e9 / peak},
                         'step_alg_bytes_per_atom': b1 + b2,
                         'step_frac_of_peak': (b1 + b2) * world * n * steps / (ms * 1e-3) / 1e9 / (peak * world)},
        }
        if slab:
            out['config']['halo'] = {'planes': 2, 'bytes_per_exchange_per_side': 2 * a.ncell[0] * a.ncell[1] * 2 * (24 if lay.get('planes') else 32),
                                     'exchanges_per_step': 2, 'timeout_flag': slab_err}
        if secondary:
            out['secondary'] = secondary
        if world == 1 and not a.no_cpu:
            os.environ['OMP_NUM_THREADS'] = str(cores)
            v, cms, cn, cores, one = cpu_leg(cpu_ncell, a.solver, a.temp, a.damping, 3, 1, one_thread_steps=1)
            out['cpu_baseline'] = {'value': v, 'unit': 'atom-steps/s', 'cores': cores, 'kind': 'port',
                                   'sample': 'bcc %dx%dx%d (%d spins) x 3 full steps, restated Fortran loops (oracle/), OpenMP over '
                                             'all host cores, noise pre-generated' % (*cpu_ncell, cn),
                                   'one_thread': {'value': one, 'unit': 'atom-steps/s', 'steps': 1}}
    else:
        out = None
    # ---- secondary: ONE supercell cut into z-slabs, halo push over NVLink fused into the boundary-tile launches (N > 1) ----
    if world > 1 and not slab and not a.no_secondary and not a.full_ham:
        e.close()
        slab_block(a, world, rank, local, dist, torch, out, warmup, steps)
        return
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        e.close()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
