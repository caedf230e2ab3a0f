import sys, time
sys.path.insert(0, '.')
from blangsdk_b200 import datasets as D
from blangsdk_b200.engine import Engine
R = 592
eng = Engine(nChains=8, nReplicas=R, random=1); eng.set_model(D.beta_binomial_batch(R)); eng.initialize("FORWARD")
eng.run_scans(5, want=())
c0 = eng.counters(); t=time.time(); eng.run_scans(20, want=()); dt=time.time()-t; c=eng.counters()
print('C5-592 ms/scan %.3f evals/s %.0f evals/exec %.2f' % (dt*50, (c['evals']-c0['evals'])/dt, (c['evals']-c0['evals'])/(c['executions']-c0['executions'])))
eng.close()
