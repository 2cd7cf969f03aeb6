#!/usr/bin/env python
"""Throughput of BASELINE config 3: FeCo B2 (tests/FeCo tables: 2 sublattices, z = 258, maptype 2), Metropolis Monte
Carlo with Mensemble = 8, plus the LLG step on the same system.  Development tool."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from oracle import inputs, orc
    from uppasd_b200 import host
    from util import GOLDEN
    nc = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    mens = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    fx = json.load(open(os.path.join(GOLDEN, 'feco.json')))
    args = list(inputs.load_fixture(fx))
    args[0] = dict(args[0], ncell=(nc, nc, nc), mensemble=mens)
    t0 = time.perf_counter()
    S = orc.build_system(*args)
    n = S['Natom']
    print('FeCo B2 %d^3: %d atoms x %d ensembles, z = %d, nHam = %d (host tables built in %.1f s)'
          % (nc, n, mens, S['exchange']['z'], S['nHam'], time.perf_counter() - t0), flush=True)
    e = host.engine_from_system(S, orc.CONST, sdealgh=1, delta_t=1e-16, damping=0.1, temp=600.0,
                                lattice_hint=(S['NA'], (nc, nc, nc), args[0]['bc']))
    t0 = time.perf_counter()
    e.mc_sweeps('M', 1, 600.0)
    e.synchronize()
    lay, ncol, per = e.mc_colouring()
    print('MC layout built in %.2f s: %d colours' % (time.perf_counter() - t0, ncol), flush=True)
    for mode in ('M', 'H'):
        e.mc_sweeps(mode, 3, 600.0)
        ms = e.time_mc_sweeps(mode, 20, 600.0)
        rate = n * mens * 20 / (ms * 1e-3)
        balg = 56 + 4 * S['exchange']['z']
        print('FeCo MC %s | %.3f ms/sweep | %.3e attempts/s | roof(%d B) %.3f' % (mode, ms / 20, rate, balg, balg * rate / 6550.1e9), flush=True)
    e.sd_steps(3)
    ms = e.time_sd_steps(20, first_step=4)
    rate = n * mens * 20 / (ms * 1e-3)
    balg = 136 + 8 * S['exchange']['z']
    print('FeCo LLG midpoint 600 K | %s | %.3f ms/step | %.3e atom-steps/s | roof(%d B) %.3f'
          % (e.layout_info(), ms / 20, rate, balg, balg * rate / 6550.1e9), flush=True)


if __name__ == '__main__':
    main()
