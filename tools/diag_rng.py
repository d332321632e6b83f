import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'oracle')); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np
import lt_oracle as O, util
from lens_trace_b200 import capi, layouts as L
ctx=capi.Context(0)
rng=np.random.default_rng(1)
n=2_000_00
W,H=1920,1080
ix=rng.integers(0,W,n); iy=rng.integers(0,H,n)
fx=(ix.astype(np.float32)/np.float32(W)+np.float32(-0.5)).astype(np.float32)
fy=(iy.astype(np.float32)/np.float32(H)+np.float32(-0.5)).astype(np.float32)
seed=rng.integers(0,3000,n).astype(np.float32)
dev=ctx.debug_random(fx,fy,seed)
host=np.array([O.random(float(a),float(b),float(c)) for a,b,c in zip(fx,fy,seed)],dtype=np.float32)
bad=np.nonzero(dev.view(np.uint32)!=host.view(np.uint32))[0]
print('random mismatches',len(bad),'of',n)
for i in bad[:10]:
    d=np.float32(np.float32(fx[i]*np.float32(12.9898))+np.float32(fy[i]*np.float32(78.233)))
    x=np.float64(d)+1113.1*np.float64(seed[i])
    print(i,fx[i],fy[i],seed[i],'dev',dev[i],'host',host[i],'x',x,'fmod',np.fmod(x,np.pi), 'a', np.sin(np.fmod(x,np.pi))*43758.5453)
sb=util.scene('cornell_box'); sc=ctx.upload(sb)
cam=util.default_camera(0.0,0)
got=ctx.render(sc,cam,capi.make_params(L.KERNEL_ACCUMULATOR,128,96))
want=O.render(L.KERNEL_ACCUMULATOR,sb,cam,128,96,threads=0)
bad=np.argwhere((got.view(np.uint32)!=want.view(np.uint32)).any(-1))
print('pixels',len(bad))
for y,x in bad[:12]:
    print(y,x,got[y,x],want[y,x])
