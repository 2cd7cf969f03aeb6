"""Time-dependent global field of the measurement phase: the magnetic-field pulse `do_bpulse` 1-4 (host side, numpy).

  pulse shapes / file   source/Fields/fieldpulse.f90:34-119 (bpulse, squarepulse, exppulse, gaussianpulse, polexppulse), :178-212 (read_bpulse)
  when it is evaluated  source/sd_driver.f90:389-393 (before the first step, at t = delta_t * rstep), :770-779 (after the moment update of
                        step mstep, at t = delta_t * mstep, every bpulse_step-th step)
  where it enters       calc_external_time_fields (calculatefields.f90:143-150) -> time_external_field -> beff2 (hamiltonianactions.f90:241)

The schedule -- one vector per step -- is handed to the engine once (asd_set_time_field); the stage kernels add the vector of their step.
"""
import math

import numpy as np


def read_bpulse(path, do_bpulse):
    """bpulsefile: a title line, b0(3), dt, step, npar, then npar parameters; returns dict(b0, step, par, ba, bb)"""
    with open(path) as fh:
        rows = [l.split() for l in fh.read().splitlines()[1:] if l.split()]
    num = lambda t: float(t.replace('d', 'e').replace('D', 'e'))
    b0 = [num(x) for x in rows[0][:3]]
    step, npar = int(num(rows[2][0])), int(num(rows[3][0]))
    par = [num(rows[4 + i][0]) for i in range(npar)] + [0.0] * (10 - npar)
    ba = bb = 0.0
    if do_bpulse == 1:
        ba = (par[0] - par[1]) / math.log(par[4] / par[5])
        bb = (par[2] - par[3]) / math.log(par[4] / par[5])
    elif do_bpulse == 2:
        ba, bb = 1.0, -1.0 / (2.0 * par[2] ** 2)
    elif do_bpulse == 3:
        ba = par[2] / (par[1] - par[0])
        bb = 1.0 / ((par[1] - par[0]) ** par[2] * math.exp(-ba * par[1]))
    return dict(b0=b0, step=step, par=par, ba=ba, bb=bb)


def pulse_amplitude(do_bpulse, P, t):
    par, ba, bb = P['par'], P['ba'], P['bb']
    if do_bpulse == 1:                                   # exppulse: plateau with exponential head and tail
        if t <= par[1]:
            return par[5] * math.exp((t - par[1]) / ba)
        if t < par[2]:
            return par[5]
        return par[5] * math.exp((par[2] - t) / bb)
    if do_bpulse == 2:                                   # gaussianpulse
        return par[5] * ba * math.exp(bb * (t - par[1]) ** 2)
    if do_bpulse == 3:                                   # polexppulse
        return par[5] * bb * (t - par[0]) ** par[2] * math.exp(-ba * t)
    if do_bpulse == 4:                                   # squarepulse
        return par[5] if par[1] <= t < par[2] else 0.0
    return 0.0


def bpulse_schedule(do_bpulse, P, delta_t, rstep, nstep):
    """tfield(3, nstep): the pulse field seen by the steps mstep = rstep + 1 ... rstep + nstep of sd_mphase"""
    out = np.zeros((3, nstep), order='F')
    amp = pulse_amplitude(do_bpulse, P, delta_t * rstep)
    scount = 1
    for s in range(nstep):
        mstep = rstep + 1 + s
        out[:, s] = [P['b0'][0] * amp, P['b0'][1] * amp, P['b0'][2] * amp]
        if scount == P['step']:
            amp = pulse_amplitude(do_bpulse, P, delta_t * mstep)
            scount = 1
        else:
            scount += 1
    return out
