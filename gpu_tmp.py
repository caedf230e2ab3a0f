import sys, time, os, numpy as np
sys.path.insert(0, '.')
from blangsdk_b200 import datasets as D
from blangsdk_b200.engine import Engine, microbench
spec = D.logistic_regression()
print('terms/s', microbench(1))
for thr in (512, 768):
  os.environ['BLANGPT_LOGISTIC_THREADS'] = str(thr)
  for G in (2,):
    eng = Engine(nChains=64, random=1, ctasPerChain=G); eng.set_model(spec); eng.initialize("COPIES")
    eng.run_scans(4, want=())
    c0 = eng.counters()['evals']; t=time.time(); out = eng.run_scans(10, want=("energy",)); dt=time.time()-t; c=eng.counters()['evals']
    print('threads', thr, 'G', G, 'ms/scan %.1f' % (dt*100), 'Mevals/s %.3f' % ((c-c0)/dt/1e6), 'E0 %.1f' % out['energy'][-1,0,0])
    eng.close()
