"""Numerics study of the inverse-factor formulation (CPU only, test infrastructure: uses the oracle). Not collected by pytest; run: python tests/numerics_study.py"""
import numpy as np, sys, scipy.linalg as sl
sys.path.insert(0,'.')
from oracle import gp_oracle as O
rng = np.random.default_rng(0)
D=8; M=1024; N=400
X = rng.standard_normal((N,D)); Zall = rng.standard_normal((M,D))
for lam, ll in [(1e-5, np.log(1/D)), (1e-10, np.log(1/D)), (1e-5, np.log(1.0)), (1e-5, np.log(4.0))]:
    g = lambda a,b: O.ard_gram(0.0, np.full(D,ll), a, b)
    Kzz = g(Zall,Zall); A = O.add_diagonal_regulariser(Kzz, lam, False)
    ev = np.linalg.eigvalsh(A); cond = ev[-1]/ev[0]
    Kxz = g(X,Zall)
    # truth in longdouble
    Al = A.astype(np.longdouble); 
    # longdouble cholesky by hand (slow but M=1024 ok?) use mp-free approach: iterative refinement of solve in longdouble
    cf = sl.cho_factor(A)
    sol = sl.cho_solve(cf, Kxz.T)
    solL = sol.astype(np.longdouble)
    for _ in range(6):
        r = Kxz.T.astype(np.longdouble) - Al @ solL
        solL = solL + sl.cho_solve(cf, r.astype(np.float64)).astype(np.longdouble)
    v_true = (1.0 - np.einsum('ij,ji->i', Kxz.astype(np.longdouble), solL)).astype(np.float64)
    v_ref = 1.0 - np.einsum('ij,ji->i', Kxz, sol)
    L = np.linalg.cholesky(A)
    Linv = sl.solve_triangular(L, np.eye(M), lower=True)
    B = Kxz @ Linv.T
    v_inv = 1.0 - np.sum(B*B,axis=1)
    Bs = sl.solve_triangular(L, Kxz.T, lower=True)
    v_trsm = 1.0 - np.sum(Bs*Bs,axis=0)
    # block inverse with block size nb
    def blockinv(nb):
        b = np.zeros((M,N))
        for c in range(0,M,nb):
            u = Kxz.T[c:c+nb] - L[c:c+nb,:c] @ b[:c]
            Li = sl.solve_triangular(L[c:c+nb,c:c+nb], np.eye(nb), lower=True)
            b[c:c+nb] = Li @ u
        return 1.0 - np.sum(b*b,axis=0)
    sc = np.max(np.abs(v_true))
    print(f"lam={lam} ll={ll:.2f} cond={cond:.2e} min v={v_true.min():.3e} max v={v_true.max():.3e}")
    for name, v in [('ref cho_solve',v_ref),('explicit inv',v_inv),('trsm',v_trsm),('blockinv32',blockinv(32)),('blockinv256',blockinv(256))]:
        print(f"   {name:14s} max|dv|/max|v| vs truth {np.max(np.abs(v-v_true))/sc:.2e}   vs ref {np.max(np.abs(v-v_ref))/sc:.2e}")
