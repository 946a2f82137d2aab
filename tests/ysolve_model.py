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


def rank_correction(P, r, a_in, b_in):
    """What k3_rank_correct adds to a slab solved with ZERO incoming carries (single-pass y-slab mode):
    -r * (a_in r^(i+1) (1 - r^(2(P-i))) / (1 - r^2) + b_in r^(P-i)), with the kernel's cut-off: rows
    further than n_cut = 41.6 / -ln r (r^n < 2^-60) from both edges, taken per 32-row segment, are skipped."""
    i = np.arange(P)
    lr = np.log(r)
    d = -r * (a_in * np.exp((i + 1) * lr) * (1.0 - np.exp(2 * (P - i) * lr)) / (1.0 - r * r) + b_in * np.exp((P - i) * lr))
    ncut = min(P, int(-41.6 / lr) + 1)
    seg0 = (i // CH) * CH
    keep = (seg0 < ncut) | (P - (seg0 + CH - 1) <= ncut)
    return np.where(keep, d, 0.0)


def solve_cyclic(g, e):
    r = root(e)
    items, chunks = slab_items(g, r)
    As, Be = close_cyclic(items, 1.0 / (1.0 - r ** len(g)))
    return apply(g, r, chunks, As, Be)


# ---- the persistent kernel's formulation (k3_ysolve_pipe) -------------------------------------
def column_table(r, L=CH):
    """Rows CT_CA / CT_CB of Plan::coltab (csrc/qg_api.cu build_plan): coefficients of the forward
    carry A and the backward carry B at row i of a full chunk once both zero-carry sweeps are done,
    cA[i] = r^(i+1) * sum_{m<L-i} r^(2m),  cB[i] = r^(L-i)."""
    i = np.arange(L)
    gs = np.concatenate(([0.0], np.cumsum(r ** (2.0 * np.arange(L + 1)))))   # gs[n] = sum_{m<n} r^(2m)
    return r ** (i + 1.0) * gs[L - i], r ** (L - i + 0.0)


def chunk_double_sweep(b, r, seg=8):
    """Zero-carry forward then backward recurrence of one chunk, computed as the kernel does: `seg`-row
    segments swept independently, then a carry fix from segment to segment.  Returns
    (z_loc, F = y_loc[last], G = z_loc[first])."""
    L = len(b)
    assert L % seg == 0
    p = r ** (np.arange(seg) + 1.0)                  # p8[k] = r^(k+1)
    y = np.array(b, dtype=float)
    for s in range(0, L, seg):                       # local forward chains
        for k in range(1, seg):
            y[s + k] += r * y[s + k - 1]
    for s in range(seg, L, seg):                     # f1, f2, f3 ...: y[s-1] is by now the true end value below
        y[s:s + seg] += p * y[s - 1]
    F = y[L - 1]
    z = y.copy()
    for s in range(0, L, seg):                       # local backward chains
        for k in range(seg - 2, -1, -1):
            z[s + k] += r * z[s + k + 1]
    for s in range(L - 2 * seg, -1, -seg):           # e3, e2, e1: true first value of the segment above
        z[s:s + seg] += p[::-1] * z[s + seg]
    return z, F, z[0]


def scan_closure(items, inv1=None, a_in=None, b_in=None):
    """close_cyclic / close_open written the way the kernel evaluates them: inclusive compositions
    of the affine maps (Kogge-Stone over the lanes), exclusive prefixes, one multiply by inv1."""
    n = len(items)
    R1, T1 = [], []
    R, T = 1.0, 0.0
    for FF, RR, _, _ in items:                       # inclusive forward composition
        T = FF + RR * T
        R = RR * R
        R1.append(R); T1.append(T)
    as0 = a_in if a_in is not None else T1[-1] * inv1
    As = [as0] + [R1[i - 1] * as0 + T1[i - 1] for i in range(1, n)]
    GGp = [X + Y * a for (_, _, X, Y), a in zip(items, As)]
    R2, T2 = [0.0] * n, [0.0] * n
    R, T = 1.0, 0.0
    for i in range(n - 1, -1, -1):                   # inclusive backward composition, applied from the top
        T = GGp[i] + items[i][1] * T
        R = items[i][1] * R
        R2[i], T2[i] = R, T
    blast = b_in if b_in is not None else T2[0] * inv1
    # carry into item i from above: the maps of items i+1 .. n-1 applied to blast, i.e. the suffix
    # composition EXCLUSIVE of item i; with constant-coefficient maps this is what the kernel's
    # shuffled (Rx, Tx) = (R2, T2) of lane i+1 hold
    Be = [0.0] * n
    Be[n - 1] = blast
    for i in range(n - 2, -1, -1):
        Be[i] = GGp[i + 1] + items[i + 1][1] * Be[i + 1]
    return As, Be


def solve_cyclic_pipe(g, e, nctas=2):
    """The whole y-solve as k3_ysolve_pipe does it: chunks -> CTAs -> cluster, element-wise apply."""
    r = root(e)
    n = len(g)
    assert n % (CH * nctas) == 0
    cA, cB = column_table(r)
    chunks = [chunk_double_sweep(g[s:s + CH], r) for s in range(0, n, CH)]
    rho, h = r ** CH, r * (1 - r ** (2 * CH)) / (1 - r * r)
    per = len(chunks) // nctas
    cta_items = [[(F, rho, G, h) for _, F, G in chunks[c * per:(c + 1) * per]] for c in range(nctas)]
    cluster = [aggregate(it) for it in cta_items]
    As_c, Be_c = scan_closure(cluster, inv1=1.0 / (1.0 - r ** n))
    u = np.zeros(n)
    for c in range(nctas):
        As, Be = scan_closure(cta_items[c], a_in=As_c[c], b_in=Be_c[c])
        for k in range(per):
            z = chunks[c * per + k][0]
            s = (c * per + k) * CH
            u[s:s + CH] = -r * (z + As[k] * cA + Be[k] * cB)
    return u
