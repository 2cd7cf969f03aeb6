#!/usr/bin/env python
"""Throughput of BASELINE config 4 (2-D triangular Heisenberg + interfacial DMI + uniaxial anisotropy + field, the
skyrmion-lattice ingredients) through the run-directory driver: LLG midpoint steps and heat-bath sweeps on a
1024 x 1024 x 1 supercell, 2 ensembles.  Development tool; algorithmic bytes per atom-step:
136 + 8 z + 8 z_dm + 104 (per-site anisotropy) with z = z_dm = 6."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from test_gpu_skyrmion import _write
    from uppasd_b200 import driver
    n1, n2 = [int(x) for x in (sys.argv[1:3] if len(sys.argv) > 2 else (1024, 1024))]
    d = tempfile.mkdtemp()
    path = _write(d, ncell=(n1, n2, 1), temp=0.0)
    sim = driver.Simulation(path)
    e = sim.engine
    natom, mens = sim.natom, sim.mens
    print('SKY layout', e.layout_info(), 'atoms', natom, 'ensembles', mens, flush=True)
    balg = 136 + 8 * 6 + 8 * 6 + 104
    for temp in (0.0, 10.0):
        sim.relax('S', nstep=10, temperature=temp, timestep=1e-16, damping=0.3)
        ms = e.time_sd_steps(100, first_step=1000)
        rate = natom * mens * 100 / (ms * 1e-3)
        print('SKY LLG midpoint T=%g | %.4f ms/step | %.3e atom-steps/s | roof(%d B) %.3f'
              % (temp, ms / 100, rate, balg, balg * rate / 6550.1e9), flush=True)
    e.mc_sweeps('H', 3, 10.0)
    ms = e.time_mc_sweeps('H', 20, 10.0)
    rate = natom * mens * 20 / (ms * 1e-3)
    print('SKY heat bath T=10 | %.4f ms/sweep | %.3e attempts/s | colouring %s' % (ms / 20, rate, e.mc_colouring()), flush=True)


if __name__ == '__main__':
    main()
