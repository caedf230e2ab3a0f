import sys, time, os, numpy as np
sys.path.insert(0, '.')
from blangsdk_b200 import datasets as D
from blangsdk_b200.engine import Engine
for fam, spec, S in (("logistic", D.logistic_regression(1000, 32), 64), ("normal", D.normal_meanvar(1000), 64)):
  for G in (1, 2, 4):
    eng = Engine(nChains=64, random=1, ctasPerChain=G); eng.set_model(spec); eng.initialize("COPIES")
    eng.run_scans(4, want=())
    c0 = eng.counters(); t=time.time(); eng.run_scans(20, want=()); dt=time.time()-t; c=eng.counters()
    ev = c['evals']-c0['evals']; ex = c['executions']-c0['executions']
    print(fam, 'G', G, 'ms/scan %.3f' % (dt*50), 'evals/exec %.2f' % (ev/ex), 'us per eval per chain %.2f' % (dt*1e6/(ev/64)), 'us per exec %.2f' % (dt*1e6/(ex/64)))
    eng.close()
