import sys, time, numpy as np
sys.path.insert(0, '.')
from blangsdk_b200 import datasets as D
from blangsdk_b200.engine import Engine
from oracle import oracle as O
import __graft_entry__ as ge
ge.smoke()

def compare(spec, nChains, scans, seed=3, G=0, **kw):
    eng = Engine(nChains=nChains, random=seed, ctasPerChain=G, **kw)
    eng.set_model(spec); eng.initialize("FORWARD")
    ref = O.PT(O.Model(spec), nChains=nChains, seed=seed, nThreads=8); ref.initialize()
    d0 = eng.log_density(); r0 = ref.densities()
    print(' init loglik rel diff', np.max(np.abs(d0['loglik'][0]-r0['loglik'])/np.maximum(1,np.abs(r0['loglik']))),
          'fixed', np.max(np.abs(d0['fixed'][0]-r0['fixed'])/np.maximum(1,np.abs(r0['fixed']))), 'noos', d0['noos'][0], r0['noos'])
    t=time.time(); out = eng.run_scans(scans); tg=time.time()-t
    e_ref=[];i_ref=[]
    t=time.time()
    for _ in range(scans):
        e,ind=ref.scan(); e_ref.append(e); i_ref.append(ind)
    tc=time.time()-t
    e_ref=np.array(e_ref); i_ref=np.array(i_ref)
    same = np.array_equal(out['swapIndicators'][:,0,:], i_ref)
    rel = np.max(np.abs(out['energy'][:,0,:]-e_ref)/np.maximum(1,np.abs(e_ref)))
    c = eng.counters()
    print(' swaps same', same, 'energy rel', rel, 'evals gpu', c['evals'], 'oracle', ref.n_evals(), 'gpu s', round(tg,3), 'cpu s', round(tc,3))
    eng.close()

print('normal'); compare(D.normal_meanvar(), 8, 50)
print('logistic small'); compare(D.logistic_regression(n=2000, p=8), 8, 10)
print('logistic small G2'); compare(D.logistic_regression(n=2001, p=8), 8, 10, G=2)
print('mixture small'); compare(D.gaussian_mixture(n=3000), 8, 10)
print('mixture small G4'); compare(D.gaussian_mixture(n=3001), 8, 10, G=4)
print('betabin'); compare(D.beta_binomial(), 8, 30)
# timing C2
spec = D.logistic_regression()
for G in (1,2):
    eng = Engine(nChains=64, random=1, ctasPerChain=G); eng.set_model(spec); eng.initialize("FORWARD")
    eng.run_scans(1, want=())
    c0 = eng.counters(); t=time.time(); eng.run_scans(3, want=()); dt=time.time()-t; c1=eng.counters()
    print('C2 G',G,'scans/s', 3/dt, 'evals/s', (c1['evals']-c0['evals'])/dt)
    eng.close()
