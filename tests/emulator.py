"""numpy emulator of the device walk program (csrc/schedule.hpp op semantics, the arithmetic of
felsenstein_walk in csrc/kernel_walk.cuh) — lets the host scheduler be checked without a GPU.
All columns are processed at once as arrays; `reg` plays the per-thread register `cur`."""
import numpy as np

OPK_LEAF, OPK_REG, OPK_MEM, OPK_CHERRY = 0, 1, 2, 3
PREM_ROOT, PREM_REG, PREM_STACK = 0, 1, 2
OUT_NONE, OUT_KEEP, OUT_PUSH = 0, 1, 2


def _rescale(v):
    m = v.max(axis=0)
    _, e = np.frexp(m)           # m = f * 2**e, f in [0.5, 1)
    e = e - 1                    # bring the max into [1, 2)
    return v * np.ldexp(1.0, -e)[None, :], e


def run_program(prog, codes, K, P, dP, pi, n_real_branches, eigen=None):
    """prog: capi.schedule_dump output.  codes (rows, S) uint8.  P, dP: (K, K, R, NB) Fortran
    arrays [s_parent, s_child, r, branch] for the REAL branches.  Returns (ll, grad_dev) with
    grad_dev indexed by device branch id.

    eigen = dict(U, D, Uinv, mu, blv, rates, NE) switches INTERNAL children to the walk kernel's own
    arithmetic (csrc/device_math.cuh): P L = L + U (em1 * (Uinv L)) and P^T q = q + Uinv^T (em1 * (U^T q))
    over the first NE eigen-components only (the host has moved the null eigenvalue last,
    capi.model_reorder), and the gradient numerator in eigen-space.  As in the kernel, a post op applies
    the branch ABOVE its node before the result is carried on or stored (slots and `reg` hold D = P L), a
    stored child of the gradient pass is used as it is and its numerator is sum_i (U^T q)_i c_i (Uinv D)_i
    with c = D mu rate (dP = U diag(c) Uinv P); a recomputed cherry goes the old way, sum_i (U^T q)_i de_i (Uinv L)_i.
    Leaf children keep using the P / dP table columns, as on the device."""
    R = P.shape[2]
    S = codes.shape[1]
    nd = prog["n_dnodes"]

    def tables(br, r):
        if br < n_real_branches:
            return P[:, :, r, br], dP[:, :, r, br]
        return np.eye(K), np.zeros((K, K))

    def leaf_down(T, row):
        # column pick, all-ones -> row sums
        ext = np.concatenate([T, T.sum(axis=1, keepdims=True)], axis=1)   # (K, K+1)
        c = np.full(S, K) if row < 0 else np.minimum(codes[row].astype(int), K)
        return ext[:, c]

    dstore = eigen is not None and K >= 3          # kernel_walk.cuh: DST (slots hold D = P L) for K >= 3 only
    if eigen is not None:
        NE = eigen["NE"]
        Ue, Uie = np.asarray(eigen["U"])[:, :NE], np.asarray(eigen["Uinv"])[:NE, :]

        def eig_vecs(br, r):
            if br >= n_real_branches:            # virtual identity branch
                return np.zeros(NE), np.zeros(NE)
            x = eigen["mu"] * eigen["blv"][br] * np.asarray(eigen["D"])[:NE] * eigen["rates"][r]
            return np.expm1(x), np.asarray(eigen["D"])[:NE] * eigen["mu"] * eigen["rates"][r] * np.exp(x)

    ll = 0.0
    grad = np.zeros(nd)
    for r in range(R):
        slots = {}
        reg = np.ones((K, S))
        esum = np.zeros(S, dtype=np.int64)
        for op in prog["post"]:
            a_src, a_br, b_src, b_br, dst, flags, node = (int(v) for v in op[:7])

            def down(kind, src, br):
                Pm, _ = tables(br, r)
                if kind == OPK_LEAF:
                    return leaf_down(Pm, src)
                L = reg if kind == OPK_REG else slots[src]
                if dstore:
                    return L                     # already D = P L: the child's op applied its branch
                if eigen is not None:
                    em1, _ = eig_vecs(br, r)
                    return L + Ue @ (em1[:, None] * (Uie @ L))
                return Pm @ L
            Da = down(flags & 3, a_src, a_br)
            Db = down((flags >> 2) & 3, b_src, b_br)
            reg, e = _rescale(Da * Db)
            esum += e
            if dstore and not (flags & 32):                     # the branch above this node (the root has none)
                em1, _ = eig_vecs(node, r)
                reg = reg + Ue @ (em1[:, None] * (Uie @ reg))
            if flags & 16:
                slots[dst] = reg
        ll += (esum * np.log(2.0) + np.log(pi @ reg)).sum()

        stack = {}
        for op in prog["pre"]:
            a_src, a_br, b_src, b_br, m_src, flags, a_dst, b_dst = (int(v) for v in op)
            mk = (flags >> 8) & 3
            pm = np.repeat(pi[:, None], S, axis=1) if mk == PREM_ROOT else (reg if mk == PREM_REG else stack[m_src])

            def child(kind, src, br, dst):
                Pm, dPm = tables(br, r)
                internal = kind != OPK_LEAF
                if internal:
                    if kind == OPK_CHERRY:     # rebuilt from the two leaves below it, as the post pass built it
                        Lx = leaf_down(tables(dst & 0xffff, r)[0], src & 0xffff)
                        Ly = leaf_down(tables((dst >> 16) & 0xffff, r)[0], (src >> 16) & 0xffff)
                        L, _ = _rescale(Lx * Ly)
                    else:
                        L = slots[src]
                    if eigen is not None:    # Y = eigen-coordinates of dP L
                        if kind == OPK_CHERRY or not dstore:
                            em1, de = eig_vecs(br, r)
                            w = Uie @ L
                            return L + Ue @ (em1[:, None] * w), de[:, None] * w, Pm
                        crate = np.zeros(NE) if br >= n_real_branches else np.asarray(eigen["D"])[:NE] * eigen["mu"] * eigen["rates"][r]
                        return L, crate[:, None] * (Uie @ L), Pm          # the slot holds D = P L
                    return Pm @ L, dPm @ L, Pm
                return leaf_down(Pm, src), leaf_down(dPm, src), Pm
            Da, Ya, Pa = child(flags & 3, a_src, a_br, a_dst)
            Db, Yb, Pb = child((flags >> 2) & 3, b_src, b_br, b_dst)
            qa, qb = pm * Db, pm * Da
            den = (qa * Da).sum(axis=0)
            for internal, br, q, Y in (((flags & 3) != OPK_LEAF, a_br, qa, Ya), (((flags >> 2) & 3) != OPK_LEAF, b_br, qb, Yb)):
                num = ((Ue.T @ q) * Y).sum(axis=0) if (eigen is not None and internal) else (q * Y).sum(axis=0)
                grad[br] += (num / den).sum()
            for out, dst, Pm, q, br in (((flags >> 10) & 3, a_dst, Pa, qa, a_br), ((flags >> 12) & 3, b_dst, Pb, qb, b_br)):
                if out == OUT_NONE:
                    continue
                if eigen is not None:
                    em1, _ = eig_vecs(br, r)
                    v, _ = _rescale(q + Uie.T @ (em1[:, None] * (Ue.T @ q)))
                else:
                    v, _ = _rescale(Pm.T @ q)
                if out == OUT_KEEP:
                    reg = v
                else:
                    stack[dst] = v
    return ll, grad
