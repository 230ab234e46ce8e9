import json, sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np
import celeritas_b200 as cb, celerref
from test_gpu_field import isotropic_mix
cfg = json.load(open('data/images/cms-scale-small.json'))
del cfg['field']; cfg['field_map'] = 'data/field/cms-tiny.field.json'
problem = celerref.Problem(cfg); problem.export_image('/tmp/rz.b2img')
params = cb.Params('/tmp/rz.b2img')
ref, gpu = problem.stepper(4096), cb.Stepper(params, 4096)
prim = isotropic_mix(24, 100.0, params, seed=3)
cr, cg = ref.step(prim), gpu.step(prim)
for it in range(120):
    assert cr == cg, (it, cr, cg)
    st = ref.get('status'); act = st != 0
    for f in ('status', 'volume_id', 'post_step_action', 'num_steps'):
        assert np.array_equal(ref.get(f), gpu.get(f)), (it, f)
    assert np.array_equal(ref.get('rng'), gpu.get('rng')), it
    dd = np.abs(ref.get('dir') - gpu.get('dir'))[act]; dp = np.abs(ref.get('pos') - gpu.get('pos'))[act]
    de = np.abs(ref.get('energy') - gpu.get('energy'))[act]
    if len(dd):
        i = np.unravel_index(dd.argmax(), dd.shape)[0]
        slot = np.nonzero(act)[0][i]
        print(it, 'dir %.2e pos %.2e energy %.2e' % (dd.max(), dp.max(), de.max()), 'slot', slot,
              'E %.4g' % ref.get('energy')[slot], 'nsteps', ref.get('num_steps')[slot], 'loop', ref.get('num_looping_steps')[slot], 'pid', ref.get('particle_id')[slot])
    if not (cr['alive'] or cr['queued']): break
    cr, cg = ref.step(), gpu.step()
