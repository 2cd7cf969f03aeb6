"""`python -m uppasd_b200.run [directory]` -- runs the inpsd.dat of a reference-style run directory on the GPU and
writes the reference's measurement files there (the role of the `sd` binary for the slice this path serves)."""
import os
import sys
import time


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    d = os.path.abspath(argv[0] if argv else '.')
    from . import driver
    t0 = time.time()
    sim = driver.Simulation(os.path.join(d, 'inpsd.dat'))
    print('uppasd_b200: %d atoms x %d ensembles, mode %s, simid %s' % (sim.natom, sim.mens, sim.inp['mode'], sim.inp['simid']))
    sim.run()
    print('uppasd_b200: done in %.2f s (%d kernel launches)' % (time.time() - t0, sim.engine.launch_count()))
    return 0


if __name__ == '__main__':
    sys.exit(main())
