#!/usr/bin/env python
"""Throughput of BASELINE config 3 as the reference's own example states it: FeCo RANDOM alloy (examples/Mappings/FeCo/random:
bcc primitive cell, 50/50 occupancy, z = 258), Mensemble = 8, Metropolis and heat-bath sweeps plus the LLG step, through the
run-directory driver (occupancy from the reference's generator, one coupling row per atom).  Development tool.
Algorithmic bytes (SURVEY 8d with per-atom couplings): 56 + 12 z per attempt, 136 + 24 z per LLG atom-step."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from uppasd_b200 import driver
    nc = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    mens = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    fx = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'feco_random.json')))
    d = tempfile.mkdtemp()
    for k, v in fx['raw'].items():
        open(os.path.join(d, k), 'w').write(v)
    drop = ('ncell', 'mensemble', 'ip_mode', 'mode', 'sdealgh', 'temp')
    lines = [l for l in fx['raw']['inpsd.dat'].splitlines() if not (l.split() and l.split()[0].lower() in drop)]
    lines += ['ncell %d %d %d' % (nc, nc, nc), 'mensemble %d' % mens, 'ip_mode N', 'mode M', 'sdealgh 1', 'temp 600']
    open(os.path.join(d, 'inpsd.dat'), 'w').write('\n'.join(lines) + '\n')
    import warnings
    warnings.simplefilter('ignore')
    t0 = time.perf_counter()
    sim = driver.Simulation(os.path.join(d, 'inpsd.dat'))
    e, n = sim.engine, sim.natom
    z = sim.tables['nlist'].shape[0]
    print('FeCo random alloy %d^3: %d atoms x %d ensembles, z = %d (set up in %.1f s)' % (nc, n, mens, z, time.perf_counter() - t0), flush=True)
    t0 = time.perf_counter()
    e.mc_sweeps('M', 1, 600.0)
    e.synchronize()
    print('MC layout built in %.2f s: %s' % (time.perf_counter() - t0, e.mc_colouring()), flush=True)
    for mode in ('M', 'H'):
        e.mc_sweeps(mode, 3, 600.0)
        ms = e.time_mc_sweeps(mode, 20, 600.0)
        rate = n * mens * 20 / (ms * 1e-3)
        balg = 56 + 12 * z
        print('alloy MC %s | %.3f ms/sweep | %.3e attempts/s | roof(%d B) %.3f' % (mode, ms / 20, rate, balg, balg * rate / 6451.2e9), flush=True)
    e.sd_steps(3)
    ms = e.time_sd_steps(20, first_step=4)
    rate = n * mens * 20 / (ms * 1e-3)
    balg = 136 + 24 * z
    print('alloy LLG midpoint | %s | %.3f ms/step | %.3e atom-steps/s | roof(%d B) %.3f'
          % (e.layout_info(), ms / 20, rate, balg, balg * rate / 6451.2e9), flush=True)


if __name__ == '__main__':
    main()
