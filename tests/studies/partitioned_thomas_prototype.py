"""NumPy prototype of the partitioned two-sided block-Thomas (S sweeps, Q = S/2 + 1 sub-domains, Q - 1 separators)."""
import numpy as np

def layout(n, S):
    Q = S // 2 + 1
    # separators split n nodes into Q sub-domains of near-equal length
    inner = n - (Q - 1)
    base, rem = divmod(inner, Q)
    lens = [base + (1 if q < rem else 0) for q in range(Q)]
    seps, starts = [], []
    pos = 0
    for q in range(Q):
        starts.append(pos)
        pos += lens[q]
        if q < Q - 1:
            seps.append(pos); pos += 1
    sweeps = []   # (first, dir, rows, start_sep, end_sep)
    for q in range(Q):
        a, b = starts[q], starts[q] + lens[q] - 1
        sL = seps[q - 1] if q > 0 else None
        sR = seps[q] if q < Q - 1 else None
        if q == 0:
            sweeps.append((a, +1, lens[q], None, sR))
        elif q == Q - 1:
            sweeps.append((b, -1, lens[q], None, sL))
        else:
            sweeps.append((a, +1, lens[q], sL, sR))      # down
            sweeps.append((b, -1, lens[q], sR, sL))      # up
    return seps, sweeps

def solve_partitioned(A, B, C, d, S):
    n, m = d.shape
    seps, sweeps = layout(n, S)
    fac = []
    for (first, dr, rows, ssep, esep) in sweeps:
        Cp = np.zeros((rows, m, m)); dp = np.zeros((rows, m)); Wp = np.zeros((rows, m, m))
        Xc = np.zeros((m, m)); Xd = np.zeros(m); Wprev = np.zeros((m, m))
        for r in range(rows):
            k = first + dr * r
            Ak = A[k] if dr > 0 else C[k]          # coupling to the node BEHIND in sweep direction
            Ck = C[k] if dr > 0 else A[k]          # coupling to the node AHEAD
            Bk = B[k].copy(); dk = d[k].copy()
            if r == 0:
                W = Ak.copy() if ssep is not None else np.zeros((m, m))
                # boundary start (ssep None): Ak couples to nothing (k = 0 or n-1)
            else:
                Bk -= Ak @ Xc; dk -= Ak @ Xd; W = -Ak @ Wprev
            inv = np.linalg.inv(Bk)
            Xc = inv @ Ck; Xd = inv @ dk; Wprev = inv @ W
            Cp[r], dp[r], Wp[r] = Xc, Xd, Wprev
        fac.append((Cp, dp, Wp))
    # separator equations
    ns = len(seps)
    L = np.zeros((ns, m, m)); D = np.zeros((ns, m, m)); R = np.zeros((ns, m, m)); rhs = np.zeros((ns, m))
    for si, s in enumerate(seps):
        D[si] = B[s]; rhs[si] = d[s]
    for (first, dr, rows, ssep, esep), (Cp, dp, Wp) in zip(sweeps, fac):
        if esep is None: continue
        si = seps.index(esep)
        last = first + dr * (rows - 1)
        Ah = A[esep] if last < esep else C[esep]   # coupling of the separator row to the sweep's last row
        D[si] -= Ah @ Cp[-1]; rhs[si] -= Ah @ dp[-1]
        if ssep is not None:
            sj = seps.index(ssep)
            if sj < si: L[si] = -Ah @ Wp[-1]
            else: R[si] = -Ah @ Wp[-1]
    # reduced block-tridiagonal solve
    xs = np.zeros((ns, m))
    Cr = np.zeros((ns, m, m)); dr_ = np.zeros((ns, m))
    for si in range(ns):
        Bk = D[si].copy(); dk = rhs[si].copy()
        if si > 0:
            Bk -= L[si] @ Cr[si - 1]; dk -= L[si] @ dr_[si - 1]
        inv = np.linalg.inv(Bk); Cr[si] = inv @ R[si]; dr_[si] = inv @ dk
    for si in range(ns - 1, -1, -1):
        xs[si] = dr_[si] - (Cr[si] @ xs[si + 1] if si + 1 < ns else 0)
    x = np.zeros((n, m))
    for si, s in enumerate(seps): x[s] = xs[si]
    # back substitution per sweep (boundary sweeps: all rows; interior: the half nearest the end separator)
    for (first, dr, rows, ssep, esep), (Cp, dp, Wp) in zip(sweeps, fac):
        xS = x[ssep] if ssep is not None else np.zeros(m)
        xnext = x[esep]
        r_stop = 0 if ssep is None else rows // 2
        for r in range(rows - 1, r_stop - 1, -1):
            k = first + dr * r
            xk = dp[r] - Cp[r] @ xnext - Wp[r] @ xS
            x[k] = xk; xnext = xk
    # interior sub-domains: rows r < rows//2 of the down sweep are covered by the up sweep's r >= rows - rows//2 ... check coverage
    return x

rng = np.random.default_rng(0)
for n in (11, 23, 64, 101, 1091):
    for S in (2, 4, 8):
        if n < 3 * S: continue
        m = 7
        A = rng.normal(size=(n, m, m)); C = rng.normal(size=(n, m, m)); B = rng.normal(size=(n, m, m)) + 8 * np.eye(m)
        A[0] = 0; C[-1] = 0
        d = rng.normal(size=(n, m))
        M = np.zeros((n * m, n * m))
        for k in range(n):
            M[k*m:(k+1)*m, k*m:(k+1)*m] = B[k]
            if k > 0: M[k*m:(k+1)*m, (k-1)*m:k*m] = A[k]
            if k < n - 1: M[k*m:(k+1)*m, (k+1)*m:(k+2)*m] = C[k]
        xref = np.linalg.solve(M, d.ravel()).reshape(n, m)
        x = solve_partitioned(A, B, C, d, S)
        cov = np.isfinite(x).all()
        print(n, S, np.abs(x - xref).max(), layout(n, S)[0][:4])
