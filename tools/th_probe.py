import sys, json
sys.path.insert(0,'.')
import numpy as np, torch
from viorb_b200 import api, synth
dev=torch.device('cuda',0); stream=torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx=api.Context(0, stream.cuda_stream)
B=512; h,w=480,752
imgs=synth.frames(64,h,w,seed0=0)
d=torch.from_numpy(np.ascontiguousarray(np.tile(imgs,(B//64,1,1)))).to(dev)
for mt in (7,12,20):
    ex=api.ORBextractor(1000,1.2,8,20,mt,ctx=ctx)
    cap=ex.cap
    dk=torch.empty((B,cap,7),dtype=torch.float32,device=dev); dd=torch.empty((B,cap,32),dtype=torch.uint8,device=dev); dc=torch.zeros((B,),dtype=torch.int32,device=dev)
    for _ in range(2): ex.extract_batch_device(d,B,h,w,dk,dd,dc)
    ex.check(); ex.profile(True)
    for _ in range(3): ex.extract_batch_device(d,B,h,w,dk,dd,dc)
    ex.check(); st,p=ex.stage_ms(); ex.profile(False)
    print('minTh',mt,{k:round(v/3,3) for k,v in st.items()}, 'kp/frame', float(dc.float().mean()))
    ex.close()
