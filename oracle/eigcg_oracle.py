"""oracle/eigcg_oracle.py -- TEST INFRASTRUCTURE (never imported by the product).

numpy restatement of the reference's eigCG and incremental eigCG:

    ks_eigCG_parity       generic_ks/inc_eigcg.c:377-850   CG for (4m^2 - D^2) x = b that also builds the Lanczos
                                                            matrix T of the same Krylov space in a window of m
                                                            vectors, restarted with the 2*Nvecs Ritz vectors of
                                                            T_m and T_{m-1} (Stathopoulos & Orginos, arXiv:0707.0131)
    initCG                inc_eigcg.c:47-121               deflated start x += U (H + 4m^2)^-1 U^+ r
    orthogonalize         inc_eigcg.c:123-213              modified Gram-Schmidt of the new vectors
    extend_H              inc_eigcg.c:215-262              H = -U^+ D^2 U grown by the new columns
    ks_inc_eigCG_parity   inc_eigcg.c:851-950
    calc_eigenpairs       inc_eigcg.c:282-300              Rayleigh-Ritz on everything accumulated

The stencil is the C oracle's (oracle/ks_oracle.c kso_dslash); the small dense problems go to numpy.linalg where
the reference calls LAPACK (eigenvectors are fixed up to a phase, the orthonormal basis of the QR step up to a
unitary: Ritz VALUES, the spanned spaces, iteration counts and solutions are what can be compared).
Pinned on the reference's own inc_eigcg.c compiled into oracle/_ref (oracle/ref_harness/eigcg_harness.c) by
tests/test_oracle.py and on the committed golden tests/golden/ref_eigcg.npz (tests/golden/make_golden_eigcg.py).
"""
import numpy as np

EVEN, ODD = 2, 1
ORTHO_EPS = 1e-15   # include/imp_ferm_links.h:409


class EigCGOracle:
    def __init__(self, oracle, dims, fat, lng):
        self.o, self.dims, self.fat, self.lng = oracle, tuple(dims), fat, lng
        self.V = int(np.prod(dims))

    # ---- parity-half complex vectors <-> MILC arrays -----------------------------------------------------
    def _sl(self, parity):
        return slice(0, self.V // 2) if parity == EVEN else slice(self.V // 2, self.V)

    def to_c(self, field, parity):
        f = field[self._sl(parity)]
        return (f[..., 0] + 1j * f[..., 1]).reshape(-1).copy()

    def to_field(self, c, parity, out=None):
        if out is None:
            out = np.zeros((self.V, 3, 2))
        v = c.reshape(-1, 3)
        out[self._sl(parity), :, 0] = v.real
        out[self._sl(parity), :, 1] = v.imag
        return out

    def dd(self, c, parity):
        """D_{p,p'} D_{p',p} c on the sites of `parity` (= Dslash^2, negative semi-definite)."""
        other = ODD if parity == EVEN else EVEN
        f = self.to_field(c, parity)
        t = self.o.dslash(self.dims, self.fat, self.lng, f, other)
        t = self.o.dslash(self.dims, self.fat, self.lng, t, parity)
        return self.to_c(t, parity)

    # ---- ks_eigCG_parity, inc_eigcg.c:377-850 ---------------------------------------------------------------
    def eigcg(self, src, dest, mass, parity, niter, max_restarts, resid, m, Nvecs, eigVec=None):
        """src, dest: complex parity-half vectors (dest = initial guess, updated in place).  eigVec: list of m
        complex vectors that receives the search space (the first Nvecs hold the Ritz vectors on return).
        Returns (iterations, eigVal[Nvecs] of -D^2, qic dict)."""
        msq_x4 = 4.0 * mass * mass
        rsqmin = resid * resid
        max_cg = max_restarts * niter
        if eigVec is None:
            eigVec = [None] * m
        qic = dict(size_r=0.0, final_rsq=0.0, final_iters=0, final_restart=0, converged=1)
        source_norm = float(np.vdot(src, src).real)
        if source_norm == 0.0:
            dest[:] = 0
            return 0, np.zeros(Nvecs), qic
        a, b, k = 1.0, 0.0, -1
        T = np.zeros((m, m), complex)      # upper triangle is what the reference's LAPACK calls read ("U")
        eigVal = np.zeros(max(2 * Nvecs, 1))
        iteration, nrestart = 0, 0
        rsq = 0.0
        ttt2 = None
        while True:
            if iteration % niter == 0 or (rsqmin <= 0 or rsqmin > qic["size_r"]):          # :520-576
                ttt = self.dd(dest, parity) - msq_x4 * dest
                resid_v = src + ttt
                cg_p = resid_v.copy()
                rsq = float(np.vdot(resid_v, resid_v).real)
                qic["final_rsq"] = rsq / source_norm
                iteration += 1
                if iteration >= max_cg or nrestart >= max_restarts or (rsqmin <= 0 or rsqmin > qic["final_rsq"]):
                    break
                if Nvecs > 0:
                    a, b, k = 1.0, 0.0, -1
                    T[:] = 0
                nrestart += 1
            ttt = self.dd(cg_p, parity) - msq_x4 * cg_p                                     # :578-597
            pkp = float(np.vdot(cg_p, ttt).real)
            if Nvecs > 0:
                if k == m - 1:                                                               # :607-690
                    Th = np.triu(T) + np.triu(T, 1).conj().T
                    w1, Y1 = np.linalg.eigh(Th)
                    w2, Y2 = np.linalg.eigh(Th[:m - 1, :m - 1])
                    Y = np.zeros((m, 2 * Nvecs), complex)
                    Y[:, :Nvecs] = Y1[:, :Nvecs]
                    Y[:m - 1, Nvecs:] = Y2[:, :Nvecs]
                    Y[m - 1, Nvecs - 1:] = 0      # :624 zeroes row m-1 of columns Nvecs-1 .. 2 Nvecs-1 (sic: one column early)
                    Q, _ = np.linalg.qr(Y)
                    Ts = Q.conj().T @ (Th @ Q)
                    Ts = (Ts + Ts.conj().T) / 2
                    ev, Z = np.linalg.eigh(Ts)
                    eigVal[:2 * Nvecs] = ev
                    QZ = Q @ Z
                    Vm = np.stack(eigVec[:m], axis=1)                  # sites x m
                    Vn = Vm @ QZ
                    for j in range(2 * Nvecs):
                        eigVec[j] = Vn[:, j].copy()
                    T[:] = 0                                            # (upper triangle: diag + the column set below)
                    for j in range(2 * Nvecs):
                        T[j, j] = ev[j]
                    k = 2 * Nvecs - 1
                    ttt2 = ttt2 - ttt
                    for j in range(2 * Nvecs):
                        T[j, k + 1] = np.vdot(eigVec[j], ttt2) / np.sqrt(rsq)
                elif k >= 0:
                    T[k, k + 1] = -np.sqrt(b) / a
                k += 1
                eigVec[k] = resid_v / np.sqrt(rsq)
                T[k, k] = b / a
            a = -rsq / pkp                                                                   # :739-776
            b = rsq
            dest += a * cg_p
            resid_v = resid_v + a * ttt
            rsq = float(np.vdot(resid_v, resid_v).real)
            iteration += 1
            qic["size_r"] = rsq / source_norm
            qic["final_iters"] = iteration
            qic["final_restart"] = nrestart
            b = rsq / b
            cg_p = resid_v + b * cg_p
            if Nvecs > 0:
                T[k, k] += 1.0 / a
                if k == m - 1:
                    ttt2 = b * ttt
        out_val = np.zeros(Nvecs)
        if Nvecs > 0:                                                                        # :795-808
            k += 1
            Th = np.triu(T[:k, :k]) + np.triu(T[:k, :k], 1).conj().T
            w, Z = np.linalg.eigh(Th)
            Vm = np.stack(eigVec[:k], axis=1)
            Vn = Vm @ Z[:, :Nvecs]
            for j in range(Nvecs):
                eigVec[j] = Vn[:, j].copy()
            out_val = w[:Nvecs] - msq_x4
        qic["final_iters"] = iteration
        qic["final_restart"] = nrestart
        qic["converged"] = 0 if (nrestart == max_restarts or iteration == max_cg) else 1
        return iteration, out_val, qic

    # ---- incremental eigCG, inc_eigcg.c:47-300,851-950 -------------------------------------------------------
    def inc_init(self, m, Nvecs, Nvecs_max):
        self.p = dict(m=m, Nvecs=Nvecs, Nvecs_curr=0, Nvecs_max=Nvecs_max)
        self.H = np.zeros((Nvecs_max, Nvecs_max), complex)
        self.vec = [None] * (Nvecs_max + m)
        self.val = np.zeros(Nvecs_max + m)

    def _init_cg(self, src, dest, mass, parity):                                             # :47-121
        n = self.p["Nvecs_curr"]
        msq_x4 = 4.0 * mass * mass
        r = src - (msq_x4 * dest - self.dd(dest, parity))
        c = np.array([np.vdot(self.vec[j], r) for j in range(n)])
        Hu = np.triu(self.H[:n, :n]) + np.triu(self.H[:n, :n], 1).conj().T + msq_x4 * np.eye(n)
        c = np.linalg.solve(Hu, c)
        for j in range(n):
            dest += c[j] * self.vec[j]

    def _orthogonalize(self, Nvecs, Nvecs_curr):                                             # :123-213
        j, add = Nvecs_curr, Nvecs
        n = Nvecs_curr + add
        while j < n:
            for k in range(j):
                self.vec[j] = self.vec[j] - np.vdot(self.vec[k], self.vec[j]) * self.vec[k]
            norm = np.sqrt(np.vdot(self.vec[j], self.vec[j]).real)
            if norm < ORTHO_EPS:
                add -= 1
                n -= 1
                for k in range(j, n):
                    self.vec[k] = self.vec[k + 1]
            else:
                self.vec[j] = self.vec[j] / norm
                j += 1
        return add

    def _extend_H(self, Nvecs, Nvecs_curr, parity):                                          # :215-262
        for j in range(Nvecs_curr, Nvecs_curr + Nvecs):
            ttt = self.dd(self.vec[j], parity)
            for k in range(Nvecs_curr + Nvecs):
                self.H[k, j] = -np.vdot(self.vec[k], ttt)

    def inc_eigcg(self, src, dest, mass, parity, niter, max_restarts, resid):               # :851-950
        p = self.p
        if p["Nvecs_curr"] == 0:
            self.H[:] = 0
        else:
            self._init_cg(src, dest, mass, parity)
        nc = p["Nvecs_curr"]
        work = self.vec[nc:nc + p["m"]]
        it, val, qic = self.eigcg(src, dest, mass, parity, niter, max_restarts, resid, p["m"], p["Nvecs"], work)
        self.vec[nc:nc + p["m"]] = work
        self.val[nc:nc + p["Nvecs"]] = val
        if p["Nvecs"] > 0:
            add = self._orthogonalize(p["Nvecs"], nc)
            self._extend_H(add, nc, parity)
            p["Nvecs_curr"] = nc + add
            p["Nvecs"] = min(p["Nvecs_max"] - p["Nvecs_curr"], p["Nvecs"])
        return it, qic

    def pairs(self):                                                                         # :282-300, 264-280
        n = self.p["Nvecs_curr"]
        Hu = np.triu(self.H[:n, :n]) + np.triu(self.H[:n, :n], 1).conj().T
        w, Z = np.linalg.eigh(Hu)
        Vm = np.stack(self.vec[:n], axis=1)
        Vn = Vm @ Z
        for j in range(n):
            self.vec[j] = Vn[:, j].copy()
        self.H[:n, :n] = np.diag(w)
        self.val[:n] = w
        return w.copy(), [self.vec[j] for j in range(n)]
