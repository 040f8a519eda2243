"""Pure-Python mirror of csrc/spline.cu (same statements) checked against scipy splrep/splev."""
import numpy as np
from scipy.interpolate import splrep, splev
import warnings

def spline_col(x, y, levels, K):
    m = len(x); K1 = K + 1; K2 = K + 2; n = m + K1; nk1 = m
    t = np.zeros(n + 2); a = np.zeros((m + 2, K1 + 2)); z = np.zeros(m + 2)
    X = lambda i: x[i - 1]
    Y = lambda i: y[i - 1]
    i = n
    for j in range(1, K1 + 1):
        t[j] = X(1); t[i] = X(m); i -= 1
    K3 = K // 2; jj = K3 + 2
    for q in range(K2, m + 1):
        t[q] = (X(jj) + X(jj - 1)) * 0.5 if K3 * 2 == K else X(jj)
        jj += 1
    def fpbspl(xv, l):
        h = [0.0] * (K1 + 1); hh = [0.0] * (K1 + 1)
        h[0] = 1.0
        for j in range(1, K + 1):
            for i in range(j): hh[i] = h[i]
            h[0] = 0.0
            for i in range(1, j + 1):
                li = l + i; lj = li - j
                if t[li] == t[lj]:
                    h[i] = 0.0
                else:
                    f = hh[i - 1] / (t[li] - t[lj])
                    h[i - 1] = h[i - 1] + f * (t[li] - xv)
                    h[i] = f * (xv - t[lj])
        return h
    l = K1
    for it in range(1, m + 1):
        xi = X(it); yi = Y(it)
        while not (xi < t[l + 1] or l == nk1): l += 1
        h = fpbspl(xi, l)
        j = l - K1
        for i in range(1, K1 + 1):
            j += 1
            piv = h[i - 1]
            if piv == 0.0: continue
            ww = a[j][1]; store = abs(piv)
            if store >= ww:
                r = ww / piv; dd = store * np.sqrt(1.0 + r * r)
            else:
                r = piv / ww; dd = ww * np.sqrt(1.0 + r * r)
            c = ww / dd; s = piv / dd; a[j][1] = dd
            s1 = yi; s2 = z[j]; z[j] = c * s2 + s * s1; yi = c * s1 - s * s2
            if i == K1: break
            i2 = 1
            for i1 in range(i + 1, K1 + 1):
                i2 += 1
                s1 = h[i1 - 1]; s2 = a[j][i2]
                a[j][i2] = c * s2 + s * s1; h[i1 - 1] = c * s1 - s * s2
    z[nk1] = z[nk1] / a[nk1][1]
    i = nk1 - 1
    for j in range(2, nk1 + 1):
        store = z[i]
        i1 = j - 1 if j <= K else K
        mm = i
        for q in range(1, i1 + 1):
            mm += 1
            store = store - z[mm] * a[i][q + 1]
        z[i] = store / a[i][1]
        i -= 1
    out = np.zeros(len(levels))
    ls = K1; l1 = K1 + 1
    for lv, arg in enumerate(levels):
        while not (arg >= t[ls] or l1 == K2):
            l1 = ls; ls -= 1
        while not (arg < t[l1] or ls == nk1):
            ls = l1; l1 = ls + 1
        h = fpbspl(arg, ls)
        sp = 0.0; ll = ls - K1
        for j in range(1, K1 + 1):
            ll += 1
            sp = sp + z[ll] * h[j - 1]
        out[lv] = sp
    return out

rng = np.random.default_rng(0)
worst = 0
for K in (1, 2, 3):
    for trial in range(40):
        m = int(rng.integers(K + 1, 40))
        x = np.sort(rng.random(m) * 20000) + np.arange(m)
        y = rng.standard_normal(m) * 10 + 0.01 * x
        lev = np.sort(rng.random(25) * 26000 - 3000)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = splev(lev, splrep(x, y, k=K))
        got = spline_col(x, y, lev, K)
        exact = np.array_equal(ref, got)
        err = np.max(np.abs(ref - got) / np.maximum(np.abs(ref), 1e-300))
        worst = max(worst, err)
        if not exact and trial < 3: print("K", K, "m", m, "not bit-equal, rel", err)
print("worst rel err", worst)
