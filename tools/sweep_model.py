"""CPU model (numpy/scipy, no GPU) behind the chain solve of the scalar asynchronous sweeps
(csrstream.cu): FGMRES(30) iterations on 7-point Poisson n^3 preconditioned by exact ILU(0) whose
triangular solves are replaced by k synchronous sweeps of
  A  plain Jacobi sweeps (every off-diagonal coupling read from the previous sweep),
  B  sweeps that solve the x-direction chain (i -> i±1) exactly, the rest Jacobi,
  C  the same for the z direction,  D  x and z exact,  E  exact triangular solves.
    python tools/sweep_model.py 96 A,B,C,D,E   ->   profiles/sweep_model_r02.log
One direction solved exactly is worth about one sweep (B3 = A4, B4 = A5)."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl, sys, time
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
N = n**3
def shift(k):  # matrix with ones at (i, i-k) for valid grid neighbours
    idx = np.arange(N)
    x = idx % n; y = (idx//n) % n; z = idx//(n*n)
    if k == 1: ok = x > 0
    elif k == n: ok = y > 0
    else: ok = z > 0
    return sp.csr_matrix((np.ones(ok.sum()), (idx[ok], idx[ok]-k)), shape=(N, N))
Sx, Sy, Sz = shift(1), shift(n), shift(n*n)
A = 6*sp.identity(N) - (Sx+Sy+Sz) - (Sx+Sy+Sz).T
A = A.tocsr()
# exact ILU(0) diag recurrence
d = np.zeros(N)
for i in range(N):
    v = 6.0
    x = i % n; y = (i//n) % n; z = i//(n*n)
    if x > 0: v -= 1.0/d[i-1]
    if y > 0: v -= 1.0/d[i-n]
    if z > 0: v -= 1.0/d[i-n*n]
    d[i] = v
Dinv = sp.diags(1.0/d)
# L = I - (Sx+Sy+Sz) Dinv (unit lower), U = D - (Sx+Sy+Sz)^T
NLx, NLy, NLz = -(Sx@Dinv), -(Sy@Dinv), -(Sz@Dinv)
NUx, NUy, NUz = -Sx.T, -Sy.T, -Sz.T
I = sp.identity(N, format='csr')
def make_prec(variant, k):
    # lower: (I + Nimp) y_new = r - Nexp y ; upper: (D + NUimp) z_new = y - NUexp z
    if variant == 'A': Li, Le, Ui, Ue = [], [NLx, NLy, NLz], [], [NUx, NUy, NUz]
    if variant == 'B': Li, Le, Ui, Ue = [NLx], [NLy, NLz], [NUx], [NUy, NUz]
    if variant == 'C': Li, Le, Ui, Ue = [NLz], [NLx, NLy], [NUz], [NUx, NUy]
    if variant == 'D': Li, Le, Ui, Ue = [NLx, NLz], [NLy], [NUx, NUz], [NUy]
    if variant == 'E': Li, Le, Ui, Ue = [NLx, NLy, NLz], [], [NUx, NUy, NUz], []
    Lm = (I + sum(Li)).tocsr() if Li else None
    Lex = sum(Le).tocsr() if Le else None
    Um = (sp.diags(d) + sum(Ui)).tocsr() if Ui else None
    Uex = sum(Ue).tocsr() if Ue else None
    def apply(r):
        y = r.copy()                         # initial guess as in the async sweeps (y0 = r)
        for _ in range(k):
            t = r - (Lex@y if Lex is not None else 0)
            y = spl.spsolve_triangular(Lm, t, lower=True) if Lm is not None else t
        zz = y/d
        for _ in range(k):
            t = y - (Uex@zz if Uex is not None else 0)
            zz = spl.spsolve_triangular(Um, t, lower=False) if Um is not None else t/d
        return zz
    return apply
def fgmres(A, b, M, tol=1e-8, restart=30, maxit=3000):
    x = np.zeros_like(b); bn = np.linalg.norm(b); its = 0
    while its < maxit:
        r = b - A@x; beta = np.linalg.norm(r)
        if beta/bn < tol: break
        V = [r/beta]; Z = []; H = np.zeros((restart+1, restart)); g = np.zeros(restart+1); g[0] = beta
        for j in range(restart):
            z = M(V[j]); Z.append(z); w = A@z
            for i in range(j+1):
                H[i, j] = w@V[i]
            for i in range(j+1):
                w = w - H[i, j]*V[i]
            H[j+1, j] = np.linalg.norm(w); V.append(w/H[j+1, j]); its += 1
            yk, res, *_ = np.linalg.lstsq(H[:j+2, :j+1], g[:j+2], rcond=None)
            rn = np.linalg.norm(H[:j+2, :j+1]@yk - g[:j+2])
            if rn/bn < tol or its >= maxit: break
        x = x + sum(yk[i]*Z[i] for i in range(len(yk)))
    return its
b = A@np.ones(N)
for variant, k in [(v, k) for v in sys.argv[2].split(',') for k in (3, 4, 5)]:
    t0 = time.time()
    it = fgmres(A, b, make_prec(variant, k))
    print(variant, k, it, round(time.time()-t0, 1), flush=True)
