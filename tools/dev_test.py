import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from gvi_gaussian_process_b200 import _lib
from oracle import gp_oracle as O
L = _lib.load(); P = _lib.ptr; chk = _lib.check
dev = 'cuda'
def T(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
def st(): return _lib.current_stream()

def factor(Z, ls, ll, lam, is_abs, log_noise=None, yz=None):
    M, D = Z.shape
    F = torch.empty(L.gvi_gp_factor_bytes(M, D)//8, dtype=torch.float64, device=dev)
    ws = torch.empty(L.gvi_gp_factor_workspace_bytes(M)//8, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    lsd = T([ls]); lld = T(np.atleast_1d(ll)); 
    ln = T([log_noise]) if log_noise is not None else None
    yzd = T(yz) if yz is not None else None
    chk(L.gvi_gp_factor_f64(P(T(Z)), M, D, P(lsd), P(lld), lld.numel(), lam, int(is_abs), P(ln), P(yzd), P(F), P(ws), P(info), st()), 'factor')
    torch.cuda.synchronize()
    return F, ws, int(info.item())

def run_case(N, M, D, lam=1e-5, ll_val=None, seed=0, check=True):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D)); Z = X[rng.permutation(N)[:M]].copy()
    ls_q, ll_q = 0.1, np.full(D, np.log(1.0/D) if ll_val is None else ll_val) + 0.05*rng.standard_normal(D)
    ls_p, ll_p = -0.2, np.full(D, np.log(1.0/D) if ll_val is None else ll_val)
    w = rng.standard_normal(D); y = np.sin(X @ w) + 0.1*rng.standard_normal(N)
    yz = np.sin(Z @ w)
    mq = 0.3*np.tanh(X @ rng.standard_normal(D))
    # gram
    Xd = T(X); Zd = T(Z)
    lsd = T([ls_q]); lld = T(ll_q)
    if check:
        n1 = min(N, 300)
        out = torch.empty(n1, M, dtype=torch.float64, device=dev)
        chk(L.gvi_ard_gram_f64(P(Xd), n1, P(Zd), M, D, P(lsd), P(lld), D, P(out), M, st()), 'gram')
        ref = O.ard_gram(ls_q, ll_q, X[:n1], Z)
        print(f"  gram max rel err {np.max(np.abs(out.cpu().numpy()-ref)/np.abs(ref)):.2e}")
    Fq, wsq, iq = factor(Z, ls_q, ll_q, lam, False)
    Fp, wsp, ip = factor(Z, ls_p, ll_p, 0.0, True, log_noise=0.0, yz=yz)
    print("  chol info", iq, ip)
    if check:
        Mp = (M+31)//32*32
        Lg = wsq[:Mp*Mp].view(Mp, Mp).cpu().numpy()[:M,:M]
        A = O.add_diagonal_regulariser(O.ard_gram(ls_q, ll_q, Z, Z), lam, False)
        Lr = np.linalg.cholesky(A)
        print(f"  chol max abs err {np.max(np.abs(np.tril(Lg)-Lr)):.2e}; cond(A) {np.linalg.cond(A):.2e}")
        Xi = wsq[Mp*Mp:2*Mp*Mp].view(Mp, Mp).cpu().numpy()[:M,:M]
        print(f"  Linv residual {np.max(np.abs(Xi@Lr-np.eye(M))):.2e}")
    mean_p = torch.empty(N, dtype=torch.float64, device=dev); var_p = torch.empty_like(mean_p); var_q = torch.empty_like(mean_p)
    pm = T(np.full(N, 0.25))
    chk(L.gvi_gp_predict_f64(P(Fq), M, D, P(Xd), N, None, None, None, P(var_q), st()), 'predict q')
    chk(L.gvi_gp_predict_f64(P(Fp), M, D, P(Xd), N, P(pm), None, P(mean_p), P(var_p), st()), 'predict p')
    torch.cuda.synchronize()
    if check:
        gq = lambda a,b: O.ard_gram(ls_q, ll_q, a, b); dq = lambda a: O.ard_gram_diag(ls_q, ll_q, a)
        gp = lambda a,b: O.ard_gram(ls_p, ll_p, a, b); dp = lambda a: O.ard_gram_diag(ls_p, ll_p, a)
        vq_ref = O.sparse_posterior_diag(gq, dq, Z, X, lam, False)
        mp_ref, vp_ref = O.exact_gp_posterior(gp, dp, np.full(N, 0.25), Z, yz, 0.0, X)
        e = lambda a, r: np.max(np.abs(a.cpu().numpy()-r))/np.max(np.abs(r))
        print(f"  var_q err {e(var_q, vq_ref):.2e}  mean_p err {e(mean_p, mp_ref):.2e}  var_p err {e(var_p, vp_ref):.2e}  min vq {vq_ref.min():.2e}")
    # fused step
    yd = T(y); mqd = T(mq); out3 = torch.zeros(3, dtype=torch.float64, device=dev)
    ws = torch.empty(L.gvi_fused_step_workspace_bytes(N)//8, dtype=torch.float64, device=dev)
    kind = _lib.REG_KINDS['renyi']
    args = (P(Fq), P(Fp), M, D, P(Xd), P(yd), P(mqd), P(pm), None, None, N, kind, 0.5, P(out3), None, None, None, P(ws), st())
    chk(L.gvi_fused_step_f64(*args), 'fused')
    torch.cuda.synchronize()
    if check:
        vq = np.clip(vq_ref, 1e-300, None)
        nll = O.gaussian_nll(y, mq, vq_ref); reg = O.projected_regularisation('renyi', mp_ref, vp_ref, mq, vq_ref, 0.5)
        o = out3.cpu().numpy()
        print(f"  fused nll {o[0]/N:.12f} ref {nll:.12f} rel {abs(o[0]/N-nll)/abs(nll):.2e}; reg {o[1]/N:.12f} ref {reg:.12f} rel {abs(o[1]/N-reg)/abs(reg):.2e}")
        # standalone risk
        rws = torch.empty(L.gvi_risk_workspace_bytes(N)//8, dtype=torch.float64, device=dev); o2 = torch.zeros(3, dtype=torch.float64, device=dev)
        chk(L.gvi_risk_f64(kind, 0.5, P(yd), P(mqd), P(var_q), P(mean_p), P(var_p), N, P(o2), P(rws), st()), 'risk')
        print("  standalone risk", (o2.cpu().numpy()[:2]/N))
    # timing
    for _ in range(2): chk(L.gvi_fused_step_f64(*args), 'fused')
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps): chk(L.gvi_fused_step_f64(*args), 'fused')
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    Mp = (M+31)//32*32
    fl = N*(2*(M*M + 2*M*D) + 2*M)
    print(f"  fused step N={N} M={M} D={D}: {ms:.2f} ms  {N/ms*1e3:.3e} points/s  {fl/ms*1e-9:.2f} TFLOP/s (algorithmic)")
    # factor timing
    torch.cuda.synchronize(); t=time.time(); factor(Z, ls_q, ll_q, lam, False); print(f"  factor time {1e3*(time.time()-t):.2f} ms")

def select_case(N, M, D, seed=0, check=True):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D)); ls, ll = 0.0, np.full(D, np.log(1.0/D))
    Xd = T(X); lsd = T([ls]); lld = T(ll)
    idx = torch.zeros(M, dtype=torch.int64, device=dev); tr = torch.zeros(M-1, dtype=torch.float64, device=dev)
    ws = torch.empty(L.gvi_select_workspace_bytes(N, M, D), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(); t = time.time()
    chk(L.gvi_cond_var_select_f64(P(Xd), N, D, M, P(lsd), P(lld), D, 1e-12, P(idx), P(tr), P(ws), st()), 'select')
    torch.cuda.synchronize(); dt = time.time()-t
    by = 4.0*N*M*M + 32.0*N*M
    print(f"  select N={N} M={M}: {dt*1e3:.1f} ms, {by/dt*1e-9:.1f} GB/s algorithmic")
    if check:
        ref = O.conditional_variance_select(X, M, lambda a,b: O.ard_gram(ls, ll, a, b), lambda a: O.ard_gram_diag(ls, ll, a), use_argsort=False)
        got = idx.cpu().numpy()
        print("  select indices equal:", np.array_equal(ref, got), "first mismatch", (np.flatnonzero(ref!=got)[:1]))

print("case small"); run_case(100, 10, 1, ll_val=np.log(10.0))
print("case mid"); run_case(5000, 200, 3, ll_val=0.0)
print("case M=1000 (pad)"); run_case(3000, 1000, 8)
print("case M=1500 NT=2"); run_case(2000, 1500, 8, ll_val=0.0)
print("case M=2048 NT=1"); run_case(3000, 2048, 16, ll_val=0.0)
print("select small"); select_case(2000, 50, 2)
print("select mid"); select_case(100000, 256, 8)
print("config3"); run_case(1000000, 1024, 8, check=False)
print("select config3-ish"); select_case(1000000, 512, 8, check=False)
