import json, sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import celeritas_b200 as cb, celerref
from parity import INT_FIELDS, REAL_FIELDS
name=sys.argv[1]; slots=int(sys.argv[2]); energy=float(sys.argv[3]); nprim=int(sys.argv[4]); iters=int(sys.argv[5])
cfg=json.load(open('data/images/%s.json'%name))
ref=celerref.Problem(cfg).stepper(slots)
params=cb.Params('data/images/%s.b2img'%name)
gpu=cb.Stepper(params,slots)
prim=cb.make_primaries(nprim, particle_id=params.find_particle(11), energy=energy,pos=(-22,0,0),direction=(1,0,0))
cr=ref.step(prim); cg=gpu.step(prim)
for it in range(iters):
    bad=False
    mx=0 if it==0 else mx
    if cr!=cg: print('step',it,'counters',cr,cg); bad=True
    st=ref.get('status'); act=st!=0
    for f in INT_FIELDS+REAL_FIELDS+['rng','dedx_range','macro_xs']:
        a=ref.get(f); b=gpu.get(f)
        if a.shape[0]==slots: a=a[act]; b=b[act]
        if a.dtype.kind=='f':
            fin=np.isfinite(a)&np.isfinite(b)
            ok=np.allclose(a[fin],b[fin],rtol=1e-7,atol=1e-7); mx=max(mx,float(np.max(np.abs(a[fin]-b[fin]),initial=0)))
        else: ok=np.array_equal(a,b)
        if not ok:
            bad=True
            idx=np.nonzero((a!=b) if a.ndim==1 else (a!=b).any(axis=1))[0]
            print('step',it,f,'nbad',len(idx),'slots',np.nonzero(act)[0][idx][:6],'ref',a[idx][:3],'gpu',b[idx][:3])
    if bad:
        for f in ['particle_id','energy','post_step_action','step_length','volume_id','energy_deposition','num_steps']:
            print(f, ref.get(f)[act][:12], gpu.get(f)[act][:12])
        break
    cr=ref.step(); cg=gpu.step()
print('done',it,'max abs real diff',mx, cr)
