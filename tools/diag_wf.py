import sys; sys.path.insert(0,'.')
import numpy as np
import sasktran2_b200 as sk
from sasktran2_b200 import scenarios
from oracle import oracle
from tests.test_gpu_parity import _add_native_probes, _oracle_wf
np.set_printoptions(linewidth=220, precision=2)
for nstr,nl,nlos,scat in [(16,25,6,True),(16,25,6,False),(16,25,3,True),(8,25,6,True),(16,12,6,True)]:
    sc = scenarios.small_wf_case(nstr=nstr, nlayers=nl, nwavel=5, nlos=nlos, interp=1, geotype=1)
    _add_native_probes(sc, scat_probe=scat)
    _,_,_,eng,atm = sk.engine_for_scenario(sc)
    res = eng.calculate_radiance(atm)
    ora, wf = _oracle_wf(oracle, sc)
    _, wfn = _oracle_wf(oracle, sc, perturb=1e-15)
    print('case',nstr,nl,nlos,scat,'rad', np.abs(res['radiance'][:,:,0]/ora['radiance']-1).max())
    for name, ref in wf.items():
        got=res[name][...,0]; scale=np.abs(ref).max(axis=0,keepdims=True)
        e=np.abs(got-ref)/scale; n=np.abs(wfn[name]-ref)/scale
        i=np.unravel_index(e.argmax(), e.shape)
        print('   %-24s err %.2e at %s  oracle-noise max %.2e (at same elem %.2e)'%(name, e.max(), i, n.max(), n[i]))
