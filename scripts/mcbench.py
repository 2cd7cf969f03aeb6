#!/usr/bin/env python
"""Monte Carlo micro-benchmark: spin-update attempts/s of the colour-parallel Metropolis / heat-bath sweeps on the bcc Fe
supercell (device-built tables).  Development tool; 256 B per attempt is SURVEY 8(d)'s algorithmic figure."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ncell = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 64, 64))]
    e, n = bench.bcc_engine(ncell, 1, 300.0, 0.5, 1, 0, 0)
    t0 = time.perf_counter()
    e.mc_sweeps('M', 1, 300.0)
    e.synchronize()
    print('MC layout (tables back to host + colouring + upload): %.2f s' % (time.perf_counter() - t0), flush=True)
    for mode in ('M', 'H'):
        e.mc_sweeps(mode, 5, 300.0)
        ms = e.time_mc_sweeps(mode, 20, 300.0)
        rate = n * 20 / (ms * 1e-3)
        print('MC %s bcc %dx%dx%d | %.3f ms/sweep | %.3e attempts/s | roof(256 B) %.3f | launches/sweep %d'
              % (mode, *ncell, ms / 20, rate, 256 * rate / 6550.1e9, 0), flush=True)
    e.close()


if __name__ == '__main__':
    main()
