"""Compare GPU workspace arrays with the host emulation of the same kernel bodies (debug tool)."""
import sys; sys.path.insert(0,'.')
import ctypes, numpy as np
import sasktran2_b200 as sk
from sasktran2_b200 import scenarios, _lib
def P(a): return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
nstr,nl,nlos=16,25,6
sc = scenarios.small_wf_case(nstr=nstr, nlayers=nl, nwavel=5, nlos=nlos, interp=1, geotype=1)
for k in list(sc.mappings):
    if k!='wf_aerosol_extinction': del sc.mappings[k]
_,_,_,eng,atm = sk.engine_for_scenario(sc)
res = eng.calculate_radiance(atm)
lib=ctypes.CDLL('tests/libhost_emul.so')
lib.emul_dump.restype=ctypes.c_longlong
nloc=sc.nloc; nw=5; G=1
d_leg=np.asfortranarray(sc.mappings['wf_aerosol_extinction']['d_legendre'][...,None])
rad=np.zeros((nw,nlos)); native=np.zeros((nw,nlos,nloc*(2+G)+1)); naz=ctypes.c_int(0)
alt=np.ascontiguousarray(sc.altitudes); cz=np.ascontiguousarray(sc.los_cos_vza); az=np.ascontiguousarray(sc.los_rel_az)
rc=lib.emul_do_radiance(nstr,nloc,nw,sc.leg_coeff.shape[0],nlos,P(alt),1,1,ctypes.c_double(sc.cos_sza),ctypes.c_double(sc.earth_radius),P(cz),P(az),P(sc.ssa),P(sc.total_extinction),P(sc.leg_coeff),P(sc.solar_irradiance),P(sc.albedo),1,P(rad),ctypes.byref(naz),P(d_leg),G,P(native))
N=nstr//2; L=nl; M=nstr
sizes={'lay_od':nw*L,'lay_secant':nw*L,'lay_trans':nw*(L+1),'lay_beta':nw*L*nstr,'kth':nw*M*L*2*N,'G':nw*M*L*4*N,'wvec':nw*M*nlos*L*2*N,'vsrc':nw*M*nlos*L,'xsol':nw*M*L*2*N,'zadj':nw*M*nlos*2*N*L,'lay_dbeta':nw*L*G*nstr,'wf_loc':nw*M*nlos*L*(G+4),'wf_src':nw*M*nlos*L,'wf_gnd':nw*nlos*3}
for name,n in sizes.items():
    a=np.zeros(n); b=np.zeros(n)
    na=_lib.lib().sk_b200_engine_debug_copy(eng._engine, name.encode(), P(a), n)
    nb=lib.emul_dump(name.encode(), P(b), n)
    if na!=n or nb!=n: print(name,'size mismatch',na,nb,n); continue
    scale=np.abs(b).max()
    d=np.abs(a-b)
    i=d.argmax()
    rel=np.abs(a-b)/np.maximum(np.abs(b),1e-300)
    print('%-10s max|d|/max|b| %.2e   max elementwise rel %.2e (|b| there %.2e)  idx %d'%(name, d.max()/scale, rel.max(), abs(b[rel.argmax()]), i))
    if name=='wf_loc':
        a5=a.reshape(nw,M,nlos,L,G+4); b5=b.reshape(nw,M,nlos,L,G+4)
        for lane,nm in enumerate(['eps','tau','om','t','s']):
            dd=np.abs(a5[...,lane]-b5[...,lane]); sc_=np.abs(b5[...,lane]).max()
            j=np.unravel_index(dd.argmax(), dd.shape)
            print('     lane %-4s max|d|/max %.2e at (w,ms,los,p)=%s  gpu %.6e emul %.6e'%(nm, dd.max()/sc_, j, a5[j+(lane,)], b5[j+(lane,)]))
