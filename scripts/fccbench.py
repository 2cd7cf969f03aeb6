#!/usr/bin/env python
"""fcc supercell (conventional cubic cell, 4 basis atoms, 2 exchange shells: z = 12 + 6 = 18), do_reduced Y, LLG midpoint:
the second lattice north_star names.  Algorithmic bytes per atom-step 136 + 8 z = 280.  Development tool."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

FCC = dict(cell=np.eye(3), bas=np.array([[0.0, 0.5, 0.5, 0.0], [0.0, 0.5, 0.0, 0.5], [0.0, 0.0, 0.5, 0.5]]), atype=np.array([1, 1, 1, 1]),
           mom=np.array([1.7, 1.7, 1.7, 1.7]), shells=np.array([[0.5, 0.5, 0.0], [1.0, 0.0, 0.0]]), J=np.array([1.2, -0.15]))


def fcc_engine(ncell, solver, temp, device=0):
    from uppasd_b200 import host, lattice
    C = bench.CONST
    ns, ca, cs, sh = lattice.stencil(FCC['cell'], FCC['bas'], FCC['atype'], np.array([2]), FCC['shells'][None], 1, np.ones((1, 2), dtype=int))
    cp = lattice.couplings(ns, ca, sh, FCC['atype'], FCC['J'][None, None, :], FCC['mom'], C['mry'], C['mub'])
    e = host.Engine(device)
    e.set_constants(C['gama'], C['k_bolt'], C['mub'], C['mry'])
    n = 4 * ncell[0] * ncell[1] * ncell[2]
    e.set_system(n, 1, 4, (np.arange(n, dtype=np.int32) % 4) + 1)
    e.build_lattice_table(0, 4, ncell, ('P', 'P', 'P'), ns, ca, cs, cp)
    e.set_llg(solver, 1e-16, landeg=1.0, lambda1=0.5, temp=temp, seed=7)
    e.commit()
    e.init_moments_tilted(0.1, FCC['mom'])
    return e, n, int(ns.max())


def main():
    ncell = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 128, 64))]
    for temp in (0.0, 300.0):
        e, n, z = fcc_engine(ncell, 1, temp)
        e.sd_steps(5)
        ms = e.time_sd_steps(50, first_step=6)
        balg = 136 + 8 * z
        print('FCC %dx%dx%d (%d spins, z = %d) T=%g | %s | %.4f ms/step | %.3e atom-steps/s | roof(%d B) %.3f'
              % (*ncell, n, z, temp, e.layout_info(), ms / 50, n * 50 / (ms * 1e-3), balg, balg * n * 50 / (ms * 1e-3) / 6550.1e9), flush=True)
        e.close()


if __name__ == '__main__':
    main()
