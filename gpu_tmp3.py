import sys, time, os
sys.path.insert(0, '.')
from blangsdk_b200 import datasets as D
from blangsdk_b200.engine import Engine
for fused in ("0", "1"):
    os.environ["BLANGPT_FUSED"] = fused
    eng = Engine(nChains=8, random=1); eng.set_model(D.normal_meanvar()); eng.initialize("FORWARD")
    eng.run_scans(50, want=())
    c0 = eng.counters()['evals']; t=time.time(); eng.run_scans(1000, want=()); dt=time.time()-t; c=eng.counters()['evals']
    print('C1 fused', fused, 'us/scan %.1f' % (dt*1e3), 'Mevals/s %.2f' % ((c-c0)/dt/1e6))
    eng.close()
