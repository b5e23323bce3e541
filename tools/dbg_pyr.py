import numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from viorb_b200 import api, synth
from oracle import oracle_py as O
img = synth.frame(480, 752, 0)
ex = api.ORBextractor(1000,1.2,8,20,7)
ex(img)
ref = O.Extractor(1000,1.2,8,20,7); ref(img)
for l in range(8):
    a = ex.pyramid(l); b = ref.pyramid(l)
    bad = np.argwhere(a != b)
    print(l, a.shape, b.shape, len(bad), bad[:6].tolist(), [(int(a[y,x]),int(b[y,x])) for y,x in bad[:6]])
    if len(bad):
        print(' rows', np.unique(bad[:,0])[:20], ' cols', np.unique(bad[:,1])[:40])
