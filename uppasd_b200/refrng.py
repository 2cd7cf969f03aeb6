"""The reference's host-side uniform generator, for the one place where a run's NUMBERS (not just its statistics) depend on
it: the random start `Initmag 1`, drawn once on the host before the first step.

  generator   source/RNG/mtprng.f90:118-138 (seeding), :191-245 (one 32-bit output), :281-293 (real in [0,1))
  random start  source/System/magnetizationinit.f90:141-178

mtprng.f90 is a Mersenne Twister whose state words are 64-bit Fortran integers and whose twist / tempering constants are
written as NEGATIVE decimal literals: the matrix constant and the two tempering masks are therefore sign-extended to 64
bits, the state words pick up high bits, and the first tempering shift (which is not masked) folds bits 32..42 of a
state word into the output.  The stream differs from a textbook MT19937 from the third output on; the reference's printed
goldens of every `Initmag 1` case (tests/Solvers, tests/Cluster, tests/HeisStripe, ...) are functions of exactly this
stream, so it is reproduced here with 64-bit two's-complement arithmetic on Python integers.
"""
import numpy as np

_W = (1 << 64) - 1
_LO32 = (1 << 32) - 1
_NSTATE, _MID = 624, 397
_UPPER, _LOWER = 1 << 31, (1 << 31) - 1
# the negative literals of mtprng.f90:199-207 as 64-bit two's-complement words
_TWIST = (-1727483681) & _W
_TEMPER_B = (-1658038656) & _W
_TEMPER_C = (-272236544) & _W


class ReferenceUniform:
    """rng_uniform of the reference when use_vsl is off: mtprng_rand_real2 on a state seeded with tseed (uppasd.f90:903-910)"""

    def __init__(self, seed):
        w = [0] * _NSTATE
        w[0] = int(seed) & _W
        for i in range(1, _NSTATE):
            prev = w[i - 1]
            w[i] = (1812433253 * (prev ^ (prev >> 30)) + i) & _LO32
        self._w, self._next = w, _NSTATE

    def _regenerate(self):
        w = self._w
        for k in range(_NSTATE):
            y = (w[k] & _UPPER) | (w[(k + 1) % _NSTATE] & _LOWER)
            w[k] = w[(k + _MID) % _NSTATE] ^ (y >> 1) ^ (_TWIST if y & 1 else 0)
        self._next = 0

    def word(self):
        if self._next >= _NSTATE:
            self._regenerate()
        y = self._w[self._next]
        self._next += 1
        y ^= y >> 11                                            # 64-bit logical shift, not masked
        y = (y ^ (((y << 7) & _W) & _TEMPER_B)) & _LO32
        y = (y ^ (((y << 15) & _W) & _TEMPER_C)) & _LO32
        return y ^ (y >> 18)

    def real2(self):
        return self.word() * (1.0 / 4294967296.0)


def random_start(na, ncell, seed):
    """Initmag 1: one direction per atom, atoms visited in storage order (cell z slowest, basis atom fastest), every
    ensemble gets the same start.  Rejection sampling in the unit ball; the reference's retry draws are scaled by 1
    instead of 2 (magnetizationinit.f90:155-160) -- kept, the goldens depend on it.  Returns emom(3, Natom)."""
    gen = ReferenceUniform(seed)
    n = na * ncell[0] * ncell[1] * ncell[2]
    out = np.empty((3, n), order='F')
    for i in range(n):
        scale = 2.0
        while True:
            x = scale * (gen.real2() - 0.50)
            y = scale * (gen.real2() - 0.50)
            z = scale * (gen.real2() - 0.50)
            if not (x * x + y * y + z * z > 1):
                break
            scale = 1.0
        nrm = np.sqrt(x * x + y * y + z * z)
        out[0, i], out[1, i], out[2, i] = x / nrm, y / nrm, z / nrm
    return out
