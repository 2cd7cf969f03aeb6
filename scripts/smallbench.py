"""Small-system stepping rate (BASELINE.json configs[0]: tests/bccFe, 6^3 bcc cells = 432 spins, T = 300 K): LLG steps per
second with the resident kernel (one launch, state in shared memory) against two stage launches per step."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from util import load_golden
from oracle import orc
from uppasd_b200 import host

NSTEP = 20000
for name in ('bccfe', 'kagome', 'cluster', 'heisstripe'):
    fx, inp, S = load_golden(name)
    for alg in (5, 1):
        for res in ('1', '0'):
            os.environ['ASD_RESIDENT'] = res
            e = host.engine_from_system(S, orc.consts(S), sdealgh=alg, delta_t=1e-16, damping=0.5, temp=300.0)
            e.sd_steps(200)
            e.synchronize()
            ms = e.time_sd_steps(NSTEP, first_step=201)
            print('%-10s N=%5d z=%3d SDEalgh %d resident=%s: %8.3f us/step  %10.0f steps/s  %.3e atom-steps/s' % (
                name, S['Natom'], S['exchange']['z'], alg, res, 1e3 * ms / NSTEP, NSTEP / ms * 1e3, S['Natom'] * NSTEP / ms * 1e3), flush=True)
            e.close()
# Monte Carlo sweeps on the same small systems (colour classes of a few dozen atoms)
for name in ('bccfe', 'kagome'):
    fx, inp, S = load_golden(name)
    for mode, res in (('M', '1'), ('M', '0'), ('H', '1'), ('H', '0')):
        os.environ['ASD_RESIDENT'] = res
        e = host.engine_from_system(S, orc.consts(S), sdealgh=1, delta_t=1e-16, damping=0.5, temp=300.0)
        e.mc_sweeps(mode, 50, 300.0)
        ms = e.time_mc_sweeps(mode, 5000, 300.0)
        lay, ncol, per = e.mc_colouring()
        print('%-10s N=%5d MC %s resident=%s: %8.3f us/sweep (%d colours)  %.3e attempts/s' % (name, S['Natom'], mode, res, 1e3 * ms / 5000, ncol,
                                                                                  S['Natom'] * 5000 / ms * 1e3), flush=True)
        e.close()
# the restated CPU path on the same system, one thread
fx, inp, S = load_golden('bccfe')
st = orc.SdState(S, 5, 1e-16, 0.5, temp=300.0)
orc.zig_setup(1)
N = S['Natom']
g = orc.fill_rngarray(3 * N).reshape((3, N, 1), order='F')
t = time.time()
for _ in range(2000):
    st.step(gauss=g)
dt = time.time() - t
print('oracle (CPU, noise pre-drawn) bccfe Depondt: %.2f us/step' % (dt / 2000 * 1e6))
