import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from util import load_golden
from test_asdio_host import materialise
from oracle import orc
from uppasd_b200 import driver
tmp = tempfile.mkdtemp()
fx, path = materialise('cluster', tmp)
sim = driver.Simulation(path)
fx, inp, S = load_golden('cluster')
orc.initmag1(S, inp['tseed'])
e = sim.engine
print('layout', e.layout_info())
for kind, key in ((0, 'exchange'), (2, 'bq')):
    lst, size, coup = e.get_table(kind)
    print(key, lst.shape, S[key]['list'].shape, np.array_equal(lst, S[key]['list']), np.array_equal(size, S[key]['listsize']),
          coup.shape, S[key]['coup'].shape, np.abs(coup - S[key]['coup']).max() if coup.shape == S[key]['coup'].shape else 'shape')
emom, emomM, mmom = e.get_moments()
print('start', np.abs(emom - S['emom']).max(), np.abs(mmom - S['mmom']).max())
b, en = e.effective_field()
rb, ren = orc.effective_field(S)
print('field', np.abs(b - rb).max() / np.abs(rb).max(), en, ren)
print('terms', e.energy_terms()[:, 0], orc.energy_terms(S)[:, 0])
print('landeg', sim.landeg[:3], S['Landeg'][:3])
sim.run_initial_phase()
cur = S
for ph in fx['ip_phases']:
    st = orc.SdState(cur, 5, ph['timestep'], ph['damping'])
    for _ in range(ph['nstep']):
        st.step()
    cur = dict(cur, emom=st.emom.copy(order='F'), emomM=st.emomM.copy(order='F'), mmom=st.mmom.copy(order='F'))
print('after ip', np.abs(e.get_moments()[0] - cur['emom']).max(), sim.inp['ip_nphase'], sim.inp['ipsdealgh'])
