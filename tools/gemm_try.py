import numpy as np, sys, time
sys.path.insert(0,'.')
import soundml_b200 as sb
from oracle import resample_oracle as R
sys.path.insert(0,'tests')
from test_resample_oracle import oracle_stages
rng=np.random.default_rng(0)
for sr,tg in [(44100,16000),(44100,48000),(48000,44100)]:
    cfg=sb.Resample.Config.create(sample_rate=sr,target=tg)
    st=oracle_stages(cfg)
    for n in (1, 500, 30000, 200000):
        x=rng.uniform(-1,1,(3,n)).astype(np.float32)
        want=R.apply_plan(x,st,cfg.l,cfg.m)
        t=time.time(); got=sb.Resample.apply(cfg.set_executor("planned"),x); dt=time.time()-t
        d=sb.Resample.apply(cfg.set_executor("direct"),x)
        pk=max(np.abs(want).max(),1e-3)
        print(sr,tg,n,got.shape,'gemm err',np.abs(got-want).max()/pk,'direct err',np.abs(d-want).max()/pk, f'{dt*1e3:.1f} ms', flush=True)
