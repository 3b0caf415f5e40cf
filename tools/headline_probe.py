import ctypes as C, os, sys, numpy as np, importlib.util
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT,"tests"))
from oracle import tspws_oracle as O
import conftest
B=conftest.load_pkg(); lib=B.load_library()
spec=importlib.util.spec_from_file_location("tgp", os.path.join(ROOT,"tests","test_gpu_parity.py")); T=importlib.util.module_from_spec(spec); spec.loader.exec_module(T)
mtr,n=8192,16384
x=T._bench_traces(mtr,n)
kw=dict(Kmax=10, unbiased=1, lrm=1)
p=O.derive_params(O.Params(**kw), n, 1.0); fr=O.frame_from_params(p,n)
m=(x.astype(np.float64).sum(axis=1)/n).astype(np.float32)
G=np.zeros((10,n)); g=O.group_index(mtr,10)
for k in range(10):
    sel=g==k; G[k]=(x[sel]-m[sel,None]).astype(np.float64).sum(axis=0)
st,ps=O.stacks(fr,G); ref=np.abs(ps)/10
# oracle with fp32-rounded group sums (isolates the cast)
st32,ps32=O.stacks(fr,G.astype(np.float32).astype(np.float64)); print("oracle on fp32-cast group sums: coh err %.2e"%np.abs(np.abs(ps32)/10-ref).max())
off=fr.coef_off
def report(tag, coh):
    err=np.abs(coh-ref); print(tag,"max err %.2e"%err.max(), "at", err.argmax())
    print("   per scale:", " ".join("%d:%.0e"%(s, err[off[s]:off[s+1]].max()) for s in range(0,fr.S,2)))
ls,ts,coh=T._plan_stack(B,lib,x,**kw); report("device fp32", coh)
lib.tspws_b200_set_option(B.OPT_FP64_TAIL,1)
ls,ts,coh=T._plan_stack(B,lib,x,**kw); report("device fp64 tail", coh)
lib.tspws_b200_set_option(B.OPT_FP64_TAIL,0)
r=T.gpu_run(B,lib,x,**kw); report("host fp32", r["coherence"])
# magnitude info
Y=O.analysis(fr,G); a=np.abs(Y); print("min |Y|/max|Y| per trace:", (a.min(axis=1)/a.max(axis=1))[:3])
i=np.abs(coh-ref).argmax(); print("coef",i,"|Y_k| there:", a[:,i], "max|Y|", a.max())
