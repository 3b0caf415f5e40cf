import numpy as np, sys
rec=np.loadtxt(sys.argv[1],dtype=np.int64)
t0=rec[:,2].min()
names={0:"colsum2",1:"post2",2:"fwd_tail",3:"combine",4:"inv2",5:"inv4",6:"finish"}
rows=[]
for kid in names:
    r=rec[rec[:,0]==kid]
    starts=np.sort(r[r[:,1]==0][:,2])
    idx=np.searchsorted(starts,r[:,2]+3000,side='right')-1
    for L in range(len(starts)):
        q=r[idx==L]
        if len(q): rows.append(((q[:,2].min()-t0)/1e3,(q[:,3].max()-t0)/1e3,names[kid],len(q),(np.median(q[:,2])-t0)/1e3))
rows.sort()
lo,hi=float(sys.argv[2]),float(sys.argv[3])
for a,b,nm,c,med in rows:
    if lo<a<hi: print(f"{a:8.1f} {b:8.1f} {b-a:6.1f} {nm} {c}  median cta start {med:8.1f}")
