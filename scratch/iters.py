import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from yanocomp_b200 import synth
from yanocomp_b200.batch import HostBatch
from yanocomp_b200.engine import Engine, EngineParams
cfg=synth.CONFIGS[2]
feats=synth.make_dataset(1,cfg['n_features'],cfg['n_pos'],cfg['reads_per_sample'],2,2)
host=HostBatch(feats,4,2); eng=Engine(EngineParams(window_size=3)); dev=eng.upload(host)
r=eng.run(dev,fetch='full')
it=r.iters
print('n',len(it),'mean',it.mean(),'max',it.max(),'pcts',np.percentile(it,[50,90,99,99.9,99.99]))
print('hist>50:',(it>50).sum(),'>100:',(it>100).sum(),'>500:',(it>500).sum(),'==1001:',(it>=1001).sum())
print('share of total E-steps by fits >50 iters:', it[it>50].sum()/it.sum())
