"""NumPy model of the y-solve algebra used by k3_ysolve*.cu (test helper, not product code).

Solves the cyclic system u[j-1] + d u[j] + u[j+1] = g[j] through the factorisation
(1 - r S^-1)(1 - r S) u = -r g with 32-row chunks, the affine chunk -> CTA -> rank carry
hierarchy and a single exchange of (FF, RR, X, Y) per level, exactly as the kernels do."""
import numpy as np

CH = 32


def root(e):
    """|r| < 1 root of r^2 + d r + 1 = 0 for d = -(2 + e), e > 0."""
    return 2.0 / ((2.0 + e) + np.sqrt(e * (e + 4.0)))


def chunk_sums(g, r):
    """Per 32-row chunk: forward end value F (zero carry), backward sum G of the zero-carry
    forward values, decay rho = r^len and the geometric factor h = sum r^(2t+1)."""
    n = len(g)
    out = []
    for s in range(0, n, CH):
        b = g[s:s + CH]
        y = 0.0
        yl = np.zeros(len(b))
        for i, v in enumerate(b):
            y = v + r * y
            yl[i] = y
        G = float(np.sum(yl * r ** np.arange(len(b))))
        L = len(b)
        out.append((y, G, r ** L, r * (1 - r ** (2 * L)) / (1 - r * r), yl))
    return out


def aggregate(items):
    """Fold a sequence of (FF, RR, X, Y) [or chunk (F, rho, G, h)] into one (FF, RR, X, Y)."""
    t, R, X, Y = 0.0, 1.0, 0.0, 0.0
    for FF, RR, Xi, Yi in items:
        X += R * (Xi + Yi * t)
        Y += R * (Yi * R)
        t = FF + RR * t
        R *= RR
    return t, R, X, Y


def close_cyclic(items, inv1):
    """Cyclic closure over a level: returns per item (carry_in_forward, carry_in_backward)."""
    n = len(items)
    t = 0.0
    for FF, RR, _, _ in items:
        t = FF + RR * t
    As = [t * inv1]
    for FF, RR, _, _ in items[:-1]:
        As.append(FF + RR * As[-1])
    GGp = [X + Y * a for (_, _, X, Y), a in zip(items, As)]
    t = 0.0
    for (FF, RR, _, _), gp in zip(reversed(items), reversed(GGp)):
        t = gp + RR * t
    Be = [0.0] * n
    Be[n - 1] = t * inv1
    for i in range(n - 1, 0, -1):
        Be[i - 1] = GGp[i] + items[i][1] * Be[i]
    return As, Be


def close_open(items, a_in, b_in):
    """Same with the end carries given (a slab inside a larger cyclic system)."""
    n = len(items)
    As = [a_in]
    for FF, RR, _, _ in items[:-1]:
        As.append(FF + RR * As[-1])
    GGp = [X + Y * a for (_, _, X, Y), a in zip(items, As)]
    Be = [0.0] * n
    Be[n - 1] = b_in
    for i in range(n - 1, 0, -1):
        Be[i - 1] = GGp[i] + items[i][1] * Be[i]
    return As, Be


def slab_items(g, r):
    """Chunk-level (FF, RR, X, Y) items of a slab: a chunk alone has X = G, Y = h."""
    return [(F, rho, G, h) for F, G, rho, h, _ in chunk_sums(g, r)], chunk_sums(g, r)


def apply(g, r, chunks, As, Be):
    u = np.zeros(len(g))
    for c, (F, G, rho, h, yl) in enumerate(chunks):
        s = c * CH
        L = len(yl)
        y = yl + As[c] * r ** (np.arange(L) + 1)
        z = Be[c]
        for i in range(L - 1, -1, -1):
            z = y[i] + r * z
            u[s + i] = -r * z
    return u


def solve_cyclic(g, e):
    r = root(e)
    items, chunks = slab_items(g, r)
    As, Be = close_cyclic(items, 1.0 / (1.0 - r ** len(g)))
    return apply(g, r, chunks, As, Be)
