#!/usr/bin/env python
"""prints asd_layout_info of bcc supercells of several shapes / boundary conditions (development tool)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from uppasd_b200 import host, lattice  # noqa: E402


def probe(ncell, bc):
    B, C = bench.BCC, bench.CONST
    ns, ca, cs, sh = lattice.stencil(B['cell'], B['bas'], B['atype'], np.array([4]), B['shells'][None], 1, np.ones((1, 4), dtype=int))
    cp = lattice.couplings(ns, ca, sh, B['atype'], B['J'][None, None, :], B['mom'], C['mry'], C['mub'])
    e = host.Engine(0)
    e.set_constants(C['gama'], C['k_bolt'], C['mub'], C['mry'])
    n = 2 * ncell[0] * ncell[1] * ncell[2]
    e.set_system(n, 1, 2, (np.arange(n, dtype=np.int32) % 2) + 1)
    e.build_lattice_table(0, 2, ncell, bc, ns, ca, cs, cp)
    e.set_llg(1, 1e-16, landeg=1.0, lambda1=0.5, temp=0.0)
    e.commit()
    print(ncell, bc, e.layout_info(), flush=True)
    e.close()


shapes = [tuple(int(x) for x in a.split('x')) for a in sys.argv[1:]] or [(64, 4, 8), (64, 6, 8), (96, 8, 8), (64, 8, 6), (64, 10, 10), (128, 8, 8)]
for nc in shapes:
    probe(nc, ('P', 'P', 'P'))
