import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from dct_b200 import api
from oracle import binding
orc=binding.load("oracle")
rng=np.random.default_rng(0)
for (H,W) in [(8,8),(16,264),(72,520),(1080,1920)]:
    px=rng.integers(0,256,size=(H,W),dtype=np.uint8)
    for q,a,lay in [(50,0,0),(95,1,1),(100,0,1)]:
        d,qc=api.dct_init(8),api.quant_init(8,q,a)
        plan=api.Plan(d,qc)
        out=plan.fwd_quant(px,lay)
        coef,var=(out if a else (out,None))
        rec=plan.dequant_idct(coef,W,H,lay,var)
        Q=orc.quant_table(q)
        wc,wv,_=orc.fwd_quant_plane(px,Q,a,lay,nthreads=4)
        wp,_=orc.dequant_idct_plane(wc,W,H,Q,a,lay,wv,nthreads=4)
        assert np.array_equal(coef,wc) and np.array_equal(rec,wp)
        off,sym=plan.rle_dev(torch.from_numpy(coef).cuda(),lay)
        plan.close()
c=api.dct_init(8); b=api.dct_forward(c,np.ones((8,8))); api.dct_free(c)
print("sanitizer workload ok")
