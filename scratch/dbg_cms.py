"""Trace the growth of real-field differences (reference vs CUDA) on the CMS-scale problem."""
import json, sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import celeritas_b200 as cb, celerref
from parity import INT_FIELDS, REAL_FIELDS
from test_gpu_field import isotropic_mix
energy=float(sys.argv[1]); nprim=int(sys.argv[2]); slots=int(sys.argv[3]); seed=int(sys.argv[4])
fuse=int(sys.argv[5]) if len(sys.argv)>5 else 0xffffffff
thresh=float(sys.argv[6]) if len(sys.argv)>6 else 1e-10
cfg=json.load(open('data/images/cms-scale-small.json'))
ref=celerref.Problem(cfg).stepper(slots)
params=cb.Params('data/images/cms-scale-small.b2img')
gpu=cb.Stepper(params,slots,fuse_threshold=fuse)
prim=isotropic_mix(nprim, energy, params, seed=seed)
cr=ref.step(prim); cg=gpu.step(prim)
watch=None; printed=0
prev={}
for it in range(100000):
    if cr!=cg: print('step',it,'counters',cr,cg); break
    st=ref.get('status'); act=st!=0
    worst=(0,None,None)
    for f in REAL_FIELDS:
        a=ref.get(f); b=gpu.get(f)
        d=np.abs(a-b); d=np.where(np.isfinite(d),d,0)
        if d.ndim>1: d=d.max(axis=1)
        d=np.where(act,d,0)
        s=int(np.argmax(d))
        if d[s]>worst[0]: worst=(float(d[s]),f,s)
    for f in INT_FIELDS:
        a=ref.get(f); b=gpu.get(f)
        if f in ('material_id','event_id','track_id','parent_id','particle_id','num_steps'):
            a=np.where(act,a,0); b=np.where(act,b,0)
        if not np.array_equal(a,b):
            print('step',it,'INT field',f,'differs at slots',np.nonzero(a!=b)[0][:8]); printed=99
    print('it',it,'worst %.3g %s slot %s'%worst, 'particle',ref.get('particle_id')[worst[2]] if worst[2] is not None else '', 'E',ref.get('energy')[worst[2]] if worst[2] is not None else '')
    if worst[0]>thresh and printed<12:
        printed+=1
        s=worst[2]
        print('--- step',it,'worst',worst)
        for f in ['particle_id','track_id','num_steps','volume_id','surface_id','post_step_action','along_step_action','energy','step_length','pos','dir','time','energy_deposition']:
            print('   ',f,'ref',ref.get(f)[s],'gpu',gpu.get(f)[s], ('prev ref %s'%(prev.get(f)[s],)) if f in prev else '')
    prev={f:ref.get(f).copy() for f in ['energy','pos','dir','step_length','volume_id']}
    if printed>=12 or not (cr['queued'] or cr['alive']): break
    cr=ref.step(); cg=gpu.step()
print('done',it,cr)
