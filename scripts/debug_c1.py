import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.linalg as sla, torch
from gpar_b200.engine import Engine, Factor
from gpar_b200.spec import lower_terms
from oracle import gpar_oracle as O
eng = Engine()
n = 200
x = np.linspace(0, 1, n)[:, None]
xo = x[::8]
terms = [dict(type="eq", variance=1.0, cols=[0], scales=[0.1])]
X = np.vstack([xo, x]); no = len(xo)
d = np.concatenate([np.full(no, 0.1), np.zeros(n)])
y = np.sin(6 * xo[:, 0])
K = O.kernel_matrix(terms, X, X) + np.diag(d + 1e-12)
try:
    Lref = sla.cholesky(K, lower=True); print("lapack ok, min diag", np.diag(Lref).min())
except Exception as e:
    print("lapack failed", e); Lref = None
Xp = np.hstack([X, np.zeros((len(X), 1))])
fac = Factor(eng, lower_terms(terms), eng.to_device(Xp).reshape(-1), 2, eng.to_device(d), eng.to_device(y), no, n)
print("info", fac.info.cpu().numpy())
J = fac.J.cpu().numpy().reshape(fac.n, fac.ld)[:, :fac.n]
L = np.tril(J)
print("nan count", np.isnan(L).sum(), "min diag", np.nanmin(np.diag(L)))
if Lref is not None:
    print("rel err L", np.linalg.norm(L - Lref) / np.linalg.norm(Lref))
    print("recon err ours", np.abs(L @ L.T - K).max(), "lapack", np.abs(Lref @ Lref.T - K).max())
    for k in range(0, fac.n, 32):
        e = np.abs(L[k:k+32] - Lref[k:k+32]).max()
        print(k, "blockrow err %.3e" % e, "diag min %.3e" % np.diag(L)[k:k+32].min())
