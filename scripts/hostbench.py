#!/usr/bin/env python
"""The drop-in case: tables built by the HOST in the reference's layout (here by the CPU restatement, as the Fortran
program would) handed over with asd_set_exchange -- with and without the supercell shape (asd_set_lattice_hint).
bcc Fe, 4 shells, do_reduced Y, midpoint, 300 K.  Development tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ncell = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 64, 64))]
    from oracle import orc
    from uppasd_b200 import host
    S = bench.oracle_bcc(ncell)
    n = S['Natom']
    for hint in (None, (2, ncell, ('P', 'P', 'P'))):
        t0 = time.perf_counter()
        e = host.engine_from_system(S, orc.CONST, sdealgh=1, delta_t=1e-16, damping=0.5, temp=300.0, lattice_hint=hint)
        e.sd_steps(5)
        e.synchronize()
        setup = time.perf_counter() - t0
        ms = e.time_sd_steps(100, first_step=6)
        print('HOST tables bcc %dx%dx%d | hint %s | %s | setup %.2f s | %.4f ms/step | %.3e atom-steps/s | roof(536 B) %.3f'
              % (*ncell, 'yes' if hint else 'no ', e.layout_info(), setup, ms / 100, n * 100 / (ms * 1e-3),
                 536 * n * 100 / (ms * 1e-3) / 6550.1e9), flush=True)
        e.close()


if __name__ == '__main__':
    main()
