import sys, os, json
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import numpy as np
from conftest import fo_fixture_names, load_fo, model_kwargs
from pyflation_b200 import cosmomodels as c
for name in fo_fixture_names():
    g = load_fo(name)
    m = getattr(c, g["cls"])(**model_kwargs(g)); m.run(saveresults=False)
    nf = g["nfields"]; nb = 2*nf+1
    a, b = m.foystart[nb:], g["foystart"][nb:]
    nz = np.abs(b) > 0
    e_ic = float(np.max(np.abs(a[nz]-b[nz])/np.abs(b[nz])))
    y, w = m.yresult[g["rows"]], g["yrows"]
    worst = 0.0
    for off in (0, 1):
        X, Y = y[:, nb+off::2], w[:, nb+off::2]
        sc = np.nanmax(np.abs(Y), axis=1, keepdims=True)
        with np.errstate(all="ignore"):
            r = np.abs(X-Y)/sc
        worst = max(worst, float(np.nanmax(r)))
    infl = g["bgrows"] <= int(g["fotendindex"])
    A, B = m.bgmodel.yresult[g["bgrows"][infl]], g["bgyrows"][infl]
    with np.errstate(all="ignore"):
        e_bg = float(np.nanmax(np.abs(A-B)/np.maximum(np.abs(B), 1e-300)))
    print("%-26s raw foystart %.2e  raw trajectory %.2e  bg %.2e" % (name, e_ic, worst, e_bg))
