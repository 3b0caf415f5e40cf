import ctypes as C, os, sys, numpy as np, importlib.util
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
spec=importlib.util.spec_from_file_location("ts_pws_b200", os.path.join(ROOT,"ts-pws_b200","__init__.py")); B=importlib.util.module_from_spec(spec); sys.modules["ts_pws_b200"]=B; spec.loader.exec_module(B)
lib=B.load_library()
g=np.load(os.path.join(ROOT,"tests","golden","real_ech_can_60.npz"))
for fp64 in (0,):
    r=B.call_tspws_main(lib, g["x"], B.default_params(), dt=float(g["dt"]), beg=float(g["beg"]))
    n=lib.tspws_b200_get_coherence(None,0); coh=np.empty(n); lib.tspws_b200_get_coherence(coh.ctypes.data_as(C.POINTER(C.c_double)), n)
    st=g["ex1_default_st"].astype(complex); ps=g["ex1_default_ps"].astype(complex); K=int(g["ex1_default_K"])
    err=np.abs(coh-np.abs(ps)/K)
    a=np.abs(st)/np.abs(st).max()
    print("max err", err.max(), "argmax", err.argmax(), "amp there", a[err.argmax()])
    for gate in (1e-6,1e-5,1e-4,1e-3,1e-2):
        m=a>=gate
        print("gate",gate,"frac",m.mean(),"max err",err[m].max())
    # per scale
    p=r["params"]; fr=lib.tspws_b200_frame_create(p.type,p.J,p.V,16501,p.s0,p.b0,p.w0,p.uni); d=B.frame_to_dict(fr)
    off=d["coef_off"]
    for s in range(0,d["S"],4):
        sl=slice(off[s],off[s+1]); print("scale",s,"D",d["D"][s],"max err %.2e"%err[sl].max(),"max amp %.2e"%a[sl].max())
