"""Second, independent CPU statement of the reference decoders: a literal dense transcription.

TEST INFRASTRUCTURE ONLY (see oracle/ldpc_oracle.c header).  PARITY UNPINNED by reference tests.

Follows src/ECC/Code/LDPC/Reference/Orig.hs:58-98 (sum-product) and Reference/Min.hs:54-104 (min-sum)
as written: `ne` is a dense M x N matrix, the check update visits (m, n) with H[m, n] = 1 and folds over
the other ones of row m in ascending column order, the variable update is
`foldr (+) orig_lam[j] (column j of ne')` over all M rows.  Pure Python floats (IEEE double, libm
tanh/atanh) so the rounding is the reference's.  Slow: use it on small inputs to pin the C oracle.
"""
import math


def hard(x):
    return x > 0                       # ECC.Types.hard, restated at GPU/Reference.hs:59-60


def signum(x):
    return 1.0 if x > 0 else (-1.0 if x < 0 else x)


def min_prime(x, y):                   # Min.hs:89
    ax, ay = abs(x), abs(y)
    return signum(x) * signum(y) * (ax if ax <= ay else ay)


def atanh_prime(x):                    # Utils.hs:113-117
    if x == 1.0 or x == -1.0:          # atanh = +-inf
        return signum(x) * 18.714973875118524
    return math.atanh(x)


def foldr1(f, xs):
    acc = xs[-1]
    for v in reversed(xs[:-1]):
        acc = f(v, acc)
    return acc


# Fast/CachedMult.hs:25-56 -- StableDiv as a (worst, product-of-the-rest) pair
def sd_lit(x):
    return (1.0, x) if x >= 1 else (x, 1.0)


def sd_mult(p, q):
    (a, b), (c, d) = p, q
    mn, mx = (a, c) if abs(a) < abs(c) else (c, a)          # absMinMax
    return (mn, b * mx * d)


def sd_div(p, c):
    a, b = p
    return b if a == c else a * (b / c)


def ldpc_cached_mult(h_dense, max_iterations, orig_lam, trace=None):
    """Fast/CachedMult.hs:233-264 (`ldpc/arraylet-cm`) on the expanded matrix: the check update divides the row's
    StableDiv product by the edge's own tanh; the posterior is orig_lam + foldr1 (+) (column of ne')."""
    M, N = len(h_dense), len(h_dense[0])
    ones = [[n for n in range(N) if h_dense[m][n]] for m in range(M)]
    cols = [[m for m in range(M) if h_dense[m][n]] for n in range(N)]
    ne = {(m, n): 0.0 for m in range(M) for n in ones[m]}
    lam = list(orig_lam)
    n_it = 0
    while True:
        if all(sum(1 for j in ones[m] if hard(lam[j])) % 2 == 0 for m in range(M)):
            return [1 if hard(v) else 0 for v in lam], n_it, True
        if n_it >= max_iterations:
            return [1 if hard(v) else 0 for v in orig_lam], n_it, False
        t = {(m, n): math.tanh(-((lam[n] - ne[(m, n)]) / 2)) for (m, n) in ne}
        ne2 = {}
        for m in range(M):
            if not ones[m]:
                continue
            row = foldr1(sd_mult, [sd_lit(t[(m, n)]) for n in ones[m]])
            for n in ones[m]:
                ne2[(m, n)] = -2 * atanh_prime(sd_div(row, t[(m, n)]))
        lam2 = [orig_lam[j] + foldr1(lambda x, y: x + y, [ne2[(m, j)] for m in cols[j]]) if cols[j] else orig_lam[j]
                for j in range(N)]
        if trace is not None:
            trace.append((dict(ne2), list(lam2)))
        ne, lam = ne2, lam2
        n_it += 1


def ldpc(h_dense, algo, max_iterations, orig_lam, trace=None):
    """Returns (bits, iterations, converged). algo: 'spa' (Orig.hs) or 'min' (Min.hs).
    trace: optional list; gets (ne' as {(m,n): v}, lam' list) appended per iteration."""
    if algo in ("spa_cm", "cm"):
        return ldpc_cached_mult(h_dense, max_iterations, orig_lam, trace)
    M = len(h_dense)
    N = len(h_dense[0])
    ones = [[n for n in range(N) if h_dense[m][n]] for m in range(M)]      # a ! (m,j) == 1, ascending j
    ne = [[0.0] * N for _ in range(M)]                                      # orig_ne = fmap (const 0) a
    lam = list(orig_lam)
    n_it = 0
    while True:
        c_hat = [1 if hard(v) else 0 for v in lam]
        ans = [sum(c_hat[j] for j in ones[m]) % 2 for m in range(M)]        # a `multStd` colVector c_hat over GF(2)
        if all(a == 0 for a in ans):
            return [1 if hard(v) else 0 for v in lam], n_it, True
        if n_it >= max_iterations:
            return [1 if hard(v) else 0 for v in orig_lam], n_it, False      # falls back to orig_lam (Orig.hs:70)
        ne2 = [[0.0] * N for _ in range(M)]
        for m in range(M):
            for n in ones[m]:
                if algo == "spa":
                    p = 1.0                                                 # product = foldl (*) 1
                    for j in ones[m]:
                        if j != n:
                            p = p * math.tanh(-((lam[j] - ne[m][j]) / 2))
                    ne2[m][n] = -2 * atanh_prime(p)
                else:
                    xs = [-(lam[j] - ne[m][j]) for j in ones[m] if j != n]
                    ne2[m][n] = (-3 / 4) * foldr1(min_prime, xs)
        lam2 = []
        for j in range(N):
            acc = orig_lam[j]                                               # U.foldr (+) orig_lam[j] (getCol j ne')
            for m in range(M - 1, -1, -1):
                acc = ne2[m][j] + acc
            lam2.append(acc)
        if trace is not None:
            trace.append(({(m, n): ne2[m][n] for m in range(M) for n in ones[m]}, list(lam2)))
        ne, lam = ne2, lam2
        n_it += 1
