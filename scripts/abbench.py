#!/usr/bin/env python
"""A/B of development builds of the library on the headline workload (bcc Fe 128^3, midpoint, 300 K): one child process per
library (ASD_LIB), prints step / stage-1 / stage-2 times.  usage: abbench.py lib1.so lib2.so ...   (development tool)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, %r)
import bench
solver = int(os.environ.get('AB_SOLVER', '1')); temp = float(os.environ.get('AB_TEMP', '300'))
e, n = bench.bcc_engine((128, 128, 128), solver, temp, 0.5, 1, 0, 0)
e.sd_steps(5)
best = 1e9
for r in range(3):
    best = min(best, e.time_sd_steps(40, first_step=6 + 40 * r) / 40)
s1, s2 = [], []
for r in range(5):
    _, (a, b) = e.time_sd_steps(0, first_step=200 + r, stages=True)
    s1.append(a); s2.append(b)
print('AB %%-28s solver %%d T=%%3.0f | step %%.4f ms  stage1 %%.4f  stage2 %%.4f | %%.3e atom-steps/s'
      %% (os.path.basename(os.environ.get('ASD_LIB', 'default')), solver, temp, best, np.median(s1), np.median(s2), n / (best * 1e-3)), flush=True)
e.close()
''' % ROOT


def main():
    libs = sys.argv[1:] or ['']
    for rep in range(int(os.environ.get('AB_REPS', '2'))):
        for lib in libs:
            env = dict(os.environ)
            if lib:
                env['ASD_LIB'] = os.path.abspath(lib)
            subprocess.run([sys.executable, '-c', CHILD], env=env)


if __name__ == '__main__':
    main()
