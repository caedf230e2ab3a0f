import sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
from blangsdk_b200 import datasets as D
from blangsdk_b200.engine import Engine
from blangsdk_b200.distributed import tensor_from_device_ptr
eng = Engine(nChains=64, random=1, ctasPerChain=2); eng.set_model(D.logistic_regression()); eng.initialize("COPIES")
eng.run_scans(5, want=())
tab = tensor_from_device_ptr(eng.L.blangpt_table_ptr(eng.h, 0), 64*4, torch.device('cuda',0)).view(64,4)
for s in range(4):
    eng.run_scans(1, want=())
    ev = tab[:,3].cpu().numpy()
    print('scan', s, 'evals per slot: min %d mean %.0f max %d  max/mean %.3f' % (ev.min(), ev.mean(), ev.max(), ev.max()/ev.mean()), 'sorted tail', np.sort(ev)[-5:])
eng.close()
