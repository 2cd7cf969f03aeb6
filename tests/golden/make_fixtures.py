#!/usr/bin/env python
"""Generates tests/golden/*.json from the reference's own regression fixtures.

Run in the build container (where /root/reference is mounted):  python tests/golden/make_fixtures.py
Each JSON holds (a) the reference fixture's inputs, token for token (inpsd.dat keywords, posfile, momfile,
jfile, dmfile rows), and (b) the expected values the reference's own test YAML pins for it
(tests/regulartests.yaml, tests/regressionResaro.yaml, tests/cudatests.yaml), with the YAML file:line cited.
The GPU box has no /root/reference; tests read only these JSON files.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle import inputs  # noqa: E402

REF = '/root/reference/tests'


def fixture(dirname, inpname='inpsd.dat', extra_inp=None, subst=None):
    d = os.path.join(REF, dirname)
    if subst:
        # templated input (the reference's runtest.sh seds a placeholder): read a substituted copy placed next to the data files
        import shutil
        import tempfile
        tmp = tempfile.mkdtemp()
        for f in os.listdir(d):
            shutil.copy(os.path.join(d, f), tmp)
        txt = open(os.path.join(d, inpname)).read()
        for a, b in subst.items():
            txt = txt.replace(a, b)
        inpname = 'inpsd.dat'
        open(os.path.join(tmp, inpname), 'w').write(txt)
        d = tmp
    inp = inputs.read_inpsd(os.path.join(d, inpname))
    if extra_inp:
        inp.update(extra_inp)
    files = inp.pop('files')
    out = {'source': os.path.normpath('tests/%s/%s' % (dirname, inpname))}
    keymap = {'posfile': 'posfile', 'momfile': 'momfile', 'exchange': 'jfile', 'dm': 'dmfile', 'bq': 'bqfile',
              'anisotropy': 'kfile'}
    for k, v in files.items():
        out[keymap[k]] = inputs._rows(v)
    # the same files verbatim, so that the product's own readers (uppasd_b200/asdio.py) can be run on a copy of the
    # reference's run directory materialised from this JSON
    raw = {'inpsd.dat': open(os.path.join(d, inpname)).read()}
    for k, v in files.items():
        raw[os.path.relpath(v, d)] = open(v).read()
    out['raw'] = raw
    inp['cell'] = [list(map(float, r)) for r in inp['cell']]
    out['inp'] = inp
    return out


def main():
    fx = {}
    # --- tests/kagome: midpoint + DMI + reduced Hamiltonian, T=0 (regulartests.yaml:132-156), tol 1e-8 abs
    f = fixture('kagome')
    f['inp']['traj_step'] = 100
    f['expected'] = {
        'averages': {'13000': [1.72521729, 1.00850955, -0.0656709236, 1.99944464]},
        'trajectory': {'atom': 2, '2400': [0.861611883, 0.507363131, -0.014408889, 2.0]},
        'tol': 1e-8, 'yaml': 'tests/regulartests.yaml:132-156'}
    fx['kagome'] = f
    # --- tests/Regression megaTest: 2-type cell, sym 1, maptype 2, hfield, T=0 (regressionResaro.yaml), 1e-8 abs
    f = fixture('Regression')
    f['expected'] = {
        'averages': {'11000': [0.145574347, 0.252312246, 1.87753739, 1.9]},
        'moment': {'atom': 128, '11000': [0.0766180766, 0.132795919, 0.988177572]},
        'coord': {'atom': 128, 'row': [3.5, 3.5, 3.51], 'type': 2, 'numb': 2},
        # field-level golden: precession + damping torque e x B + e x (e x B) of atom 128, B = the field of the step's SECOND
        # evaluation (prn_fields.f90:493-557 called from measure() before the next step); printed with es12.4
        'torques': {'atom': 128, '11000': [-0.048638, -0.049279, 0.010393, 0.070015]},
        'cumulants': {'211': [1.89999979, 3.6099992, 13.0320942, 0.666666666, 3.77522802e-31, 0.0, -6.10001753, -6.09996425]},
        'projavgs': {'11000': {'2': [2.5, 0.0, 0.191531371, 0.331974475, 2.47044706]}},
        'totenergy': {'10900': {'tot': -6.10005366, 'exc': -6.1, 'ext': -5.36620122e-05}},
        'tol': 1e-8, 'yaml': 'tests/regressionResaro.yaml:18-27,64-73,112-122,145-154,168-179,231-252'}
    fx['megatest'] = f
    # --- tests/FeCo: B2 two sublattices, sym 0, maptype 2, alpha=0.001 (regulartests.yaml:219-242), sloppy 2e-2
    f = fixture('FeCo')
    f['expected'] = {
        'averages_M': {'4700': 1.97771406},
        'cumulants': {'221': [1.86049127, 3.46770589, 12.1090509, 0.664336331]},
        'tol_abs': 2e-2, 'tol_rel': 2e-2, 'yaml': 'tests/regulartests.yaml:219-242'}
    fx['feco'] = f
    # --- tests/bccFe_cuda: reference CUDA path => Depondt actually runs (cudatests.yaml:24-46), sloppy
    f = fixture('bccFe_cuda', extra_inp={'sdealgh': 5})
    f['expected'] = {
        'averages_M': {'1700': 2.16117297},
        'cumulants': {'41': [2.09307165, 4.38839661, 19.3810564, 0.664537137]},
        'tol_abs': 2e-2, 'tol_rel': 2e-2, 'yaml': 'tests/cudatests.yaml:24-46',
        'note': 'gpu_mode 1: the reference CUDA path always integrates with Depondt (cudaMdSimulation.cu:319)'}
    fx['bccfe_cuda'] = f
    # --- tests/FeCo_cuda: same inputs as FeCo on the reference CUDA path (cudatests.yaml:48-70), sloppy
    f = fixture('FeCo_cuda', extra_inp={'sdealgh': 5})
    f['expected'] = {
        'averages_M': {'4700': 1.97771406},
        'cumulants': {'221': [1.86058404, 3.46805484, 12.1115283, 0.664335215]},
        'tol_abs': 2e-2, 'tol_rel': 2e-2, 'yaml': 'tests/cudatests.yaml:48-70'}
    fx['feco_cuda'] = f
    # --- tests/bccFe (inputs only; thermal goldens pin the reference's own RNG stream => statistical targets)
    f = fixture('bccFe', inpname='inpsd.dat.base')
    f['expected'] = {'note': 'thermal (T=500 K, tseed 5); these goldens pin the reference MT+Ziggurat stream and are '
                             'statistical targets for the GPU path', 'yaml': 'tests/regulartests.yaml:244-347',
                     'S_averages_2700': [1.69336666, -0.0338763749, 0.61172735, 1.8007911],
                     'S_cumulants_41': [1.7980927, 3.23410613, 10.4720508, 0.666264851, 0.000377665603, 1.10522259],
                     'M_cumulants_41': [1.76174871, 3.10487416, 9.65398874, 0.666191396, 0.000434917008, 0.98593137]}
    fx['bccfe'] = f
    # --- tests/kagome_cuda: TENSORIAL exchange (do_jtensor 1, jfile.tensor), random start (Initmag 1), do_reduced N, on the
    #     reference CUDA path => Depondt (cudatests.yaml:1-23), 1e-8 abs
    f = fixture('kagome_cuda', extra_inp={'sdealgh': 5})
    f['expected'] = {
        'averages': {'1300': [0.000376432611, 0.00431880575, -0.00287422668, 0.00520143861]},
        'cumulants': {'171': [0.00576048117, 3.39286723e-05, 1.28745915e-09, 0.627197795]},
        'tol': 1e-8, 'yaml': 'tests/cudatests.yaml:1-23',
        'note': 'gpu_mode 1: the reference CUDA path always integrates with Depondt (cudaMdSimulation.cu:319)'}
    fx['kagome_cuda'] = f
    # --- tests/Solvers: 100-spin chain, RANDOM start (Initmag 1, tseed 1: the reference's MT variant + rejection loop),
    #     T=0, damping 1, dt 1e-15, midpoint and Depondt (regulartests.yaml:349-385), 1e-8 abs
    f = fixture('Solvers', inpname='inpsd.dat.base', subst={'SOLVER': '1'})
    f['source'] = 'tests/Solvers/inpsd.dat.base (SOLVER -> 1 | 5 as in runtest.sh)'
    f['expected'] = {
        'averages': {'1': {'8000': [-0.136652492, 0.0180191027, -0.260168482, 0.294425255]},
                     '5': {'8000': [-0.13635776, 0.018097742, -0.260185475, 0.294308424]}},
        'tol': 1e-8, 'yaml': 'tests/regulartests.yaml:349-385'}
    fx['solvers'] = f
    # --- tests/HeisChain: start from a restart file, THERMAL SD initial phase (20000 midpoint steps at 0.1 K, damping 4:
    #     the reference's MT + Ziggurat stream through rannum), then T=0, damping 0 (regulartests.yaml:2-26), 1e-8 abs
    f = fixture('HeisChain')
    f['restart'] = inputs._rows(os.path.join(REF, 'HeisChain', 'startmom'))
    f['ip_phase'] = {'nstep': 20000, 'temp': 0.1, 'timestep': 1e-16, 'damping': 4.0,
                     'source': 'tests/HeisChain/inpsd.dat:18-20 (ip_mode S, ip_nphase 1)'}
    f['expected'] = {
        'averages': {'1000': [0.0125212146, 0.0243286358, 0.998725532, 0.999100271]},
        'totenergy': {'1400': -4.99936339},
        'tol': 1e-8, 'yaml': 'tests/regulartests.yaml:2-26'}
    fx['heischain'] = f
    # --- tests/Cluster: 43-atom finite cluster (BC 0 0 0), BIQUADRATIC exchange, anisotropy type 7 (uniaxial + cubic x ratio),
    #     random start, Depondt through two T = 0 initial phases and 30000 undamped steps (regulartests.yaml:182-205), 1e-8 abs
    f = fixture('Cluster')
    f['ip_phases'] = [{'nstep': 1000, 'temp': 0.0, 'timestep': 2e-16, 'damping': 0.10},
                      {'nstep': 4000, 'temp': 0.0, 'timestep': 1e-16, 'damping': 0.01}]
    f['ip_source'] = 'tests/Cluster/inpsd.dat:18-21 (ip_mode S, ip_nphase 2; ipSDEalgh defaults to SDEalgh = 5)'
    f['expected'] = {
        'averages': {'25000': [0.0507836321, 0.00124796445, -0.0415898147, 0.0656524743]},
        'totenergy': {'15000': {'tot': -4.04082234, 'exc': -3.39086666, 'ani': 0.00237801922, 'bq': -0.652333705}},
        'tol': 1e-8, 'yaml': 'tests/regulartests.yaml:182-205'}
    fx['cluster'] = f
    # --- tests/HeisStripe: 10 x 1 x 100 stripe (BC 0 0 P), UNIAXIAL anisotropy (type 1), random start, midpoint, T = 0
    #     (ip_mode N: no initial phase) (regulartests.yaml:105-129), 1e-8 abs
    f = fixture('HeisStripe')
    f['expected'] = {
        'averages': {'1190': [0.0267646384, 0.034416429, -0.00313586401, 0.0437112125]},
        'totenergy': {'1460': {'tot': -1.19366393, 'exc': -1.18719839, 'ani': -0.00646554288}},
        'tol': 1e-8, 'yaml': 'tests/regulartests.yaml:105-129'}
    fx['heisstripe'] = f
    # --- tests/HeisChainAF: antiferromagnetic chain, 2 atom types, random start, midpoint, T = 0; sublattice-projected
    #     averages (regulartests.yaml:54-103), 1e-8 abs
    f = fixture('HeisChainAF')
    f['expected'] = {
        'averages': {'1000': [0.0538882967, 0.0142337295, 0.00361204926, 0.0558533301]},
        'projavgs': {'5000': {'1': [0.47725269, -0.383636422, -0.225160187, -0.172904932],
                              '2': [0.930466673, 0.750387815, 0.438670217, 0.332046379]}},
        'totenergy': {'1500': {'tot': -1.53826977}},
        'tol': 1e-8, 'yaml': 'tests/regulartests.yaml:54-103'}
    fx['heischainaf'] = f
    # --- tests/SCsurf: 16 x 16 monolayer, Heisenberg + DM + uniaxial anisotropy in ATOMIC UNITS (aunits Y: every constant 1),
    #     maptype 2, random start, midpoint, T = 0 (regulartests.yaml:158-170), 1e-8 abs
    f = fixture('SCsurf')
    f['expected'] = {
        'averages': {'800': [-8.05438058e-05, -6.4699084e-05, -6.60907331e-05, 0.000122642819]},
        'tol': 1e-8, 'yaml': 'tests/regulartests.yaml:158-170'}
    fx['scsurf'] = f
    # --- random alloys (do_ralloy 1; BASELINE config 3).  The reference's test tree has no random-alloy case (no golden output):
    #     the INPUTS come from its examples, the expected values are properties (species counts, symmetric couplings) and the
    #     agreement of the product with the oracle's restatement of setup_chemicaldata / setup_neighbour_hamiltonian.
    #     examples/Mappings/RandomAlloy: bcc, two sites with Fe(80)Co(20) / Fe(20)Co(80)-like occupancy, 2 shells, sym 1
    fx['randomalloy'] = fixture('../examples/Mappings/RandomAlloy')
    #     examples/Mappings/FeCo/random: bcc primitive cell, one site, 50/50, z = 258 (the FeCo couplings of tests/FeCo)
    fx['feco_random'] = fixture('../examples/Mappings/FeCo/random')
    for k, v in fx.items():
        with open(os.path.join(HERE, k + '.json'), 'w') as fh:
            json.dump(v, fh, indent=1, default=lambda o: list(o))
        print('wrote', k)


if __name__ == '__main__':
    main()
