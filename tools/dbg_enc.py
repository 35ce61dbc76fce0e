import numpy as np, sys
sys.path.insert(0,'/root/repo')
from ro_map_b200 import core
from oracle import mon_oracle as orc
cfg=core.default_config(); ocfg=orc.default_config()
rng=np.random.default_rng(11)
n_grid=core.param_counts(cfg)[1]
grid=orc.f2h(rng.uniform(-1,1,n_grid).astype(np.float32))
pts=rng.random((4096,3),dtype=np.float32)
got=core.stage_encode(cfg,grid,pts); want=orc.encode(ocfg,grid,pts)
bad=np.argwhere(got!=want)
print("mismatch", len(bad), "of", got.size)
lv=np.bincount(bad[:,1]//2, minlength=16); print("by level", lv)
idx,w=orc.encode_corners(ocfg,pts)
for p,c in bad[:12]:
    print(p,c,pts[p], hex(got[p,c]),hex(want[p,c]), orc.h2f(got[p:p+1,c])[0], orc.h2f(want[p:p+1,c])[0], w[p,c//2])
