"""Helpers shared by the test modules: load a golden fixture into an oracle system."""
import json
import os

from oracle import inputs, orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    with open(os.path.join(GOLDEN, name + '.json')) as fh:
        fx = json.load(fh)
    args = inputs.load_fixture(fx)
    inp = args[0]
    S = orc.build_system(*args)
    return fx, inp, S
