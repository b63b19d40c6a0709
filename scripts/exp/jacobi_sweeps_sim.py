import numpy as np, torch, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import et_oracle as O
obs,pred=O.synthetic_trajectories(200000,seed=0)
st=O.norm_params(obs)
def gram(x):
    M=O.normalize(x.double(),*[s.double() for s in st]).reshape(x.shape[0],-1).numpy()
    return M.T@M
def jacobi(G,k,rel=1e-12,delta=None,maxsweeps=30):
    A=0.5*(G+G.T); m=A.shape[0]; V=np.eye(m)
    floor2=(1e-18*np.abs(np.diag(A)).max())**2
    counts=[]
    for sweep in range(maxsweeps):
        dk=np.sort(np.diag(A))[::-1][k-1]
        nrot=0
        for step in range(m-1):
            def player(pos): return 0 if pos==0 else 1+(pos-1+step)%(m-1)
            rots=[]
            for pi in range(m//2):
                p,q=player(pi),player(m-1-pi)
                if p>q:p,q=q,p
                app,aqq,apq=A[p,p],A[q,q],A[p,q]
                if apq*apq>floor2 and apq*apq>rel*rel*abs(app*aqq) and not (delta is not None and max(app,aqq)<delta*dk):
                    o=2*apq; dd=aqq-app; h=np.sqrt(dd*dd+o*o); t=o/(dd+(h if dd>=0 else -h)); c=1/np.sqrt(t*t+1); s=t*c
                    rots.append((p,q,c,s))
            nrot+=len(rots)
            for p,q,c,s in rots:
                ap,aq=A[:,p].copy(),A[:,q].copy(); A[:,p]=c*ap-s*aq; A[:,q]=s*ap+c*aq
                vp,vq=V[:,p].copy(),V[:,q].copy(); V[:,p]=c*vp-s*vq; V[:,q]=s*vp+c*vq
            for p,q,c,s in rots:
                ap,aq=A[p,:].copy(),A[q,:].copy(); A[p,:]=c*ap-s*aq; A[q,:]=s*ap+c*aq
        counts.append(nrot)
        if nrot==0:break
    d=np.diag(A); order=np.argsort(-d,kind='stable')
    return V[:,order[:k]],np.sqrt(np.maximum(d[order[:k]],0)),counts
for name,x in (('obs',obs),('pred',pred)):
    G=gram(x)
    w,v=np.linalg.eigh(G); Ut=v[:,::-1][:,:6]
    print(name,'eig ratios',(w[::-1]/w[-1])[:10])
    for delta in (None,1e-3,1e-6):
        U,S,counts=jacobi(G,6,delta=delta)
        P=U@U.T; Pt=Ut@Ut.T
        print(' delta',delta,'sweeps',len(counts),'rot/sweep',counts,'projector err %.2e'%np.linalg.norm(P-Pt),'S rel %.2e'%np.abs(S-np.sqrt(w[::-1][:6])).max())
