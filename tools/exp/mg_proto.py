import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from cases import *
from oracle.bp_oracle_cwrap import COracle
nxy = int(sys.argv[1]) if len(sys.argv)>1 else 60
m = ismip_model('C', L=20000.*nxy/260, n=(nxy,nxy,5))   # same dx as the bench config
P = oracle_problem(m); co = COracle(P)
u = np.full(P.n_dofs, 3e-16)
for k in range(3):   # a few Newton steps to get a representative J
    F,Jv,_ = co.assemble(u); ip,ix = P.pattern(); J = sp.csr_matrix((Jv,ix,ip),shape=(P.n_dofs,)*2)
    dx = spla.spsolve(J.tocsc(), F) if P.n_dofs < 60000 else spla.cg(J,F,rtol=1e-8,M=sp.diags(1/J.diagonal()))[0]
    u -= dx
F,Jv,_ = co.assemble(u); J = sp.csr_matrix((Jv,ix,ip),shape=(P.n_dofs,)*2)
n = P.n_dofs; nn = n//2
def blockdiag_inv(A):
    nb = A.shape[0]//2
    d00=A.diagonal()[0::2]; d11=A.diagonal()[1::2]
    d01=np.asarray(A[np.arange(0,2*nb,2),np.arange(1,2*nb,2)]).ravel(); d10=np.asarray(A[np.arange(1,2*nb,2),np.arange(0,2*nb,2)]).ravel()
    det=d00*d11-d01*d10
    rows=np.concatenate([np.arange(0,2*nb,2)]*2+[np.arange(1,2*nb,2)]*2); cols=np.concatenate([np.arange(0,2*nb,2),np.arange(1,2*nb,2)]*2)
    vals=np.concatenate([d11/det,-d01/det,-d10/det,d00/det])
    return sp.csr_matrix((vals,(rows,cols)),shape=A.shape)
def gersh(A,Di):
    B = abs(Di@A); return B.sum(axis=1).max()
def cheb(A,Di,lmax,r,deg,ratio):
    b=lmax; a=b/ratio; th=(b+a)/2; de=(b-a)/2; sg=th/de; rho=1/sg
    d=(Di@r)/th; z=d.copy()
    for k in range(1,deg):
        rn=1/(2*sg-rho); d=rn*rho*d+(2*rn/de)*(Di@(r-A@z)); z=z+d; rho=rn
    return z
# aggregates: level0 -> columns (same x,y)
X = m.mesh.coordinates()[m.node_vertex]
key = np.round(X[:,0]*1e3).astype(np.int64)*10**9 + np.round(X[:,1]*1e3).astype(np.int64)
_, agg0 = np.unique(key, return_inverse=True)
def P_from_agg(agg):
    nf=agg.size; nc=agg.max()+1
    rows=np.arange(2*nf); cols=2*np.repeat(agg,2)+np.tile([0,1],nf)
    return sp.csr_matrix((np.ones(2*nf),(rows,cols)),shape=(2*nf,2*nc))
def greedy_agg(A):
    nb=A.shape[0]//2
    G = sp.csr_matrix((np.ones(A.nnz),A.indices//2,A.indptr))[0::2]   # node graph from rows 0::2
    G = sp.csr_matrix(G); G.sum_duplicates()
    agg=-np.ones(nb,dtype=np.int64); na=0
    for i in range(nb):
        nbrs=G.indices[G.indptr[i]:G.indptr[i+1]]
        if agg[i]<0 and np.all(agg[nbrs]<0):
            agg[nbrs]=na; agg[i]=na; na+=1
    for i in range(nb):
        if agg[i]<0:
            nbrs=G.indices[G.indptr[i]:G.indptr[i+1]]; a=agg[nbrs]; a=a[a>=0]
            if a.size: agg[i]=a[0]
            else: agg[i]=na; na+=1
    return agg
levels=[dict(A=J)]
Pm=P_from_agg(agg0); 
while True:
    A=levels[-1]['A']; levels[-1]['P']=Pm
    Ac=(Pm.T@A@Pm).tocsr(); levels.append(dict(A=Ac))
    if Ac.shape[0]//2 < 300: break
    Pm=P_from_agg(greedy_agg(Ac))
for L in levels:
    L['Di']=blockdiag_inv(L['A']); L['lmax']=gersh(L['A'],L['Di'])
print('levels', [L['A'].shape[0]//2 for L in levels], 'nnz', [L['A'].nnz for L in levels])
coarse_lu = spla.splu(levels[-1]['A'].tocsc())
def vcycle(l, r, deg, ratio, omega):
    L=levels[l]
    if l==len(levels)-1: return coarse_lu.solve(r)
    z=cheb(L['A'],L['Di'],L['lmax'],r,deg,ratio)
    rc=L['P'].T@(r-L['A']@z)
    z=z+omega*(L['P']@vcycle(l+1,rc,deg,ratio,omega))
    z=z+cheb(L['A'],L['Di'],L['lmax'],r-L['A']@z,deg,ratio)
    return z
def pcg(A,b,M,rtol=1e-5,maxit=5000):
    x=np.zeros_like(b); r=b.copy(); z=M(r); p=z.copy(); rz=r@z; n0=np.linalg.norm(z); it=0
    while it<maxit:
        Ap=A@p; al=rz/(p@Ap); x+=al*p; r-=al*Ap; z=M(r); it+=1
        if np.linalg.norm(z)<=rtol*n0: break
        rz1=r@z; p=z+(rz1/rz)*p; rz=rz1
    return x,it
L0=levels[0]
x,it=pcg(J,F,lambda r: L0['Di']@r); print('block-jacobi its',it)
x,it=pcg(J,F,lambda r: cheb(J,L0['Di'],L0['lmax'],r,8,100.)); print('cheb(8,100) its',it,'spmv',it*8)
for deg,ratio,omega in ((2,4.,1.0),(3,6.,1.0),(3,6.,1.5),(2,4.,1.5),(3,10.,1.8)):
    t0=time.time(); x,it=pcg(J,F,lambda r: vcycle(0,r,deg,ratio,omega)); print('MG V(cheb%d,ratio %g, omega %g) its'%(deg,ratio,omega),it,'fine spmv/it', 2*deg+1, 'total fine spmv', it*(2*deg+1), '%.1fs'%(time.time()-t0))
