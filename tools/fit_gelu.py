"""Fits P(s) in  Phi(h) ~= 0.5 + 0.5 tanh(h P(h^2))  (vitk_common.cuh: gelu_fwd_bwd) against the exact-erf GELU and its
derivative: least squares, then a Nelder-Mead minimax refinement.  Prints coefficients and max abs errors per degree."""
import numpy as np
from scipy.special import erf
from scipy.optimize import least_squares, minimize
h = np.linspace(-7, 7, 28001)
Phi = 0.5*(1+erf(h/np.sqrt(2)))
phi = np.exp(-h*h/2)/np.sqrt(2*np.pi)
gelu = h*Phi
dgelu = Phi + h*phi
def model(c, h):
    s = h*h
    P = np.zeros_like(h)
    for a in c[::-1]:
        P = P*s + a
    g = h*P
    T = np.tanh(g)
    cdf = 0.5+0.5*T
    dP = np.zeros_like(h)
    for k in range(len(c)-1, 0, -1):
        dP = dP*s + k*c[k]
    gp = P + 2*s*dP
    d = cdf + 0.5*h*(1-T*T)*gp
    return h*cdf, d
for deg in (3,4,5):
    c0 = np.zeros(deg); c0[0]=np.sqrt(2/np.pi); c0[1]=np.sqrt(2/np.pi)*0.044715
    def res(c):
        y,d = model(c,h)
        return np.concatenate([(y-gelu), (d-dgelu)])*1e4
    r = least_squares(res, c0, x_scale=np.abs(c0)+1e-4)
    c = r.x
    # minimax refinement (Nelder-Mead on max error)
    f = lambda c: np.abs(res(c)).max()
    r2 = minimize(f, c, method='Nelder-Mead', options=dict(xatol=1e-12, fatol=1e-9, maxiter=40000, maxfev=40000))
    c = r2.x
    y,d = model(c,h)
    print(deg, [float('%.9g'%v) for v in c], 'max|gelu err|=%.2e max|dgelu err|=%.2e'%(np.abs(y-gelu).max(), np.abs(d-dgelu).max()))
