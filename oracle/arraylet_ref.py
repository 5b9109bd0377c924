"""Circulant-vector ("arraylet") statement of the reference decoders in pure Python floats.

TEST INFRASTRUCTURE ONLY (see oracle/ldpc_oracle.c header).  PARITY UNPINNED by reference tests.

Transcribes src/ECC/Code/LDPC/Fast/ArrayletMin.hs:23-114,167-192 (`ldpc/arraylet-min`) and
Fast/Arraylet.hs:164-186 (`ldpc/arraylet`) with their own containers: an Arraylet is (rotation, vector of sz values),
element r of the vector sitting at (row r, column (r + rotation) mod sz); a Matrixlet is a matrix of Maybe Arraylet.
Every helper below is the Haskell function of the same name.  Slow: used on a handful of frames to pin
oracle/arraylet_oracle.c, which in turn is compared with the Orig/Min-order statements.
"""
import math


def hard(x):
    return x > 0


def signum(x):
    return 1.0 if x > 0 else (-1.0 if x < 0 else x)


def min_prime(x, y):                       # ArrayletMin.hs:188-189
    ax, ay = abs(x), abs(y)
    return signum(x) * signum(y) * (ax if ax <= ay else ay)


def atanh_prime(x):                        # Utils.hs:113-117
    if x == 1.0 or x == -1.0:
        return signum(x) * 18.714973875118524
    return math.atanh(x)


def foldr1(f, xs):
    acc = xs[-1]
    for v in reversed(xs[:-1]):
        acc = f(v, acc)
    return acc


def zip_with(f):
    return lambda a, b: [f(x, y) for x, y in zip(a, b)]


# ---- Arraylet (ArrayletMin.hs:23-57) ----
def array_arraylet(sz, off, k):            # :42-43
    return (off, [k(r, (r + off) % sz) for r in range(sz)])


def fold_rows_arraylet(al):                # :45-47  (a, b) = splitAt (len - off) m ; b ++ a
    off, m = al
    cut = len(m) - off
    return m[cut:] + m[:cut]


def fold_cols_arraylet(al):                # :53-54
    return al[1]


def fold_cols_arraylet_p(k, al):           # :56-58  foldColsArraylet'
    off, m = al
    sz = len(m)
    return [k(ix, (ix + off) % sz, m[ix]) for ix in range(sz)]


# ---- Matrixlet (ArrayletMin.hs:63-114): (sz, rows of [Maybe Arraylet]) ----
def init_matrixlet(zero, sz, shifts):      # :66-77 with the power-of-two exponent already extracted (shift, -1 = 0 block)
    return (sz, [[None if s < 0 else (int(s), [zero] * sz) for s in row] for row in shifts])


def matrix_matrixlet(ml, k):               # :79-85
    sz, a = ml
    return (sz, [[None if x is None else array_arraylet(sz, x[0], lambda r_, c_, R=R, C=C: k(R * sz + r_, C * sz + c_))
                  for C, x in enumerate(row)] for R, row in enumerate(a)])


def fold_cols_matrixlet(f, ml):            # :92-96
    sz, a = ml
    out = []
    for row in a:
        out += foldr1(zip_with(f), [fold_cols_arraylet(x) for x in row if x is not None])
    return out


def fold_cols_matrixlet_p(f, g, ml):       # :98-102  foldColsMatrixlet'
    sz, a = ml
    out = []
    for R, row in enumerate(a):
        out += foldr1(zip_with(g), [fold_cols_arraylet_p(lambda r_, c_, v, R=R, C=C: f(R * sz + r_, C * sz + c_, v), x)
                                    for C, x in enumerate(row) if x is not None])
    return out


def fold_rows_matrixlet(f, ml):            # :104-108
    sz, a = ml
    out = []
    for C in range(len(a[0])):
        out += foldr1(zip_with(f), [fold_rows_arraylet(a[R][C]) for R in range(len(a)) if a[R][C] is not None])
    return out


def ldpc(sz, shifts, algo, max_iterations, orig_lam, trace=None):
    """ArrayletMin.hs:167-192 (algo 'min') / Arraylet.hs:164-186 (algo 'spa').  Returns (bits, iterations, converged).
    trace: optional list; gets (ne' Matrixlet, lam' list) appended per iteration."""
    m_let = init_matrixlet(0.0, sz, shifts)
    orig_lam = [float(v) for v in orig_lam]
    ne, lam, n = m_let, orig_lam, 0
    while True:
        ans = fold_cols_matrixlet(lambda x, y: x != y, matrix_matrixlet(m_let, lambda r, c: hard(lam[c])))
        if all(a == False for a in ans):   # noqa: E712  (U.all (== False) ans)
            return [1 if hard(v) else 0 for v in lam], n, True
        if n >= max_iterations:
            return [1 if hard(v) else 0 for v in orig_lam], n, False
        if algo == "spa":
            ne_tanh = fold_cols_matrixlet_p(lambda m, j, v: [(j, math.tanh(-((lam[j] - v) / 2)))], lambda a, b: a + b, ne)

            def upd(m, j):
                p = 1.0
                for jj, v in ne_tanh[m]:
                    if jj != j:
                        p = p * v
                return -2 * atanh_prime(p)
        else:
            ne_tanh = fold_cols_matrixlet_p(lambda m, j, v: [(j, -((lam[j] - v)))], lambda a, b: a + b, ne)

            def upd(m, j):
                return (-3 / 4) * foldr1(min_prime, [v for jj, v in ne_tanh[m] if jj != j])
        ne2 = matrix_matrixlet(ne, upd)
        lam2 = [a + b for a, b in zip(orig_lam, fold_rows_matrixlet(lambda x, y: x + y, ne2))]
        if trace is not None:
            trace.append((ne2, list(lam2)))
        ne, lam, n = ne2, lam2, n + 1


def matrixlet_to_edge_order(ml):
    """Messages of a Matrixlet in the reference edge order (row-major, ascending column) used by the traces."""
    sz, a = ml
    out = []
    for row in a:
        present = [x for x in row if x is not None]
        for r in range(sz):
            out += [x[1][r] for x in present]
    return out
