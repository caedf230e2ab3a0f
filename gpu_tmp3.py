import sys, time
sys.path.insert(0, '.')
from blangsdk_b200 import datasets as D
from blangsdk_b200.engine import Engine
spec = D.gaussian_mixture(400_000)
eng = Engine(nChains=74, random=1, ctasPerChain=2); eng.set_model(spec); eng.initialize("COPIES")
eng.run_scans(1, want=())
c0 = eng.counters()['evals']; t=time.time(); eng.run_scans(2, want=()); dt=time.time()-t; c=eng.counters()['evals']
print('C3-small ms/scan %.1f evals/s %.0f Grows/s %.1f' % (dt*500, (c-c0)/dt, (c-c0)/dt*4e5/1e9))
eng.close()
