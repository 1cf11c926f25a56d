"""NumPy stand-in for the per-panel pieces of the block-row-cyclic update (TEST INFRASTRUCTURE: lets the
world_size-2 gloo test run the distributed driver on CPU).  Same formulas as csrc/large.cuh."""

import numpy as np
import torch

from cholupdates_b200.distributed import RC_BLOCK, rc_owned_rows


class NumpyRowCyclicBackend:
    def __init__(self, A_local, n, rank, world):
        self.A = A_local.numpy()  # shares memory with the torch tensor
        self.n, self.rank, self.world = n, rank, world
        self.rows = rc_owned_rows(n, rank, world)
        self.status = torch.zeros(1, dtype=torch.int32)
        self.w = np.zeros(n)
        self.p, self.c, self.sg, self.dn = (np.zeros(n) for _ in range(4))
        self.t, self.S = 1.0, 1.0

    def new_payload(self):
        return torch.zeros(4 * RC_BLOCK + 8, dtype=torch.float64)

    def prepare(self, v, v_out, check_diag):
        self.v_out = v_out.numpy()
        vv = v.numpy()
        self.w[self.rows] = vv[self.rows]
        self.t, self.S = 1.0, 1.0
        self.status[0] = 0
        if check_diag and np.any(self.A[np.arange(len(self.rows)), self.rows] == 0):
            self.status[0] = 1

    def panel_owner(self, J, payload):
        j0, j1 = J * RC_BLOCK, min(self.n, (J + 1) * RC_BLOCK)
        loc = np.searchsorted(self.rows, np.arange(j0, j1))
        LD = np.tril(self.A[loc, j0:j1])
        p = np.linalg.solve(LD, self.w[j0:j1])  # forward substitution on the diagonal block
        dg = np.diag(LD)
        t_incl = self.t + np.cumsum(p * p)
        t_excl = np.concatenate(([self.t], t_incl[:-1]))
        sgn = np.where(dg < 0, -1.0, 1.0)
        rinv = 1.0 / np.sqrt(t_excl * t_incl)
        self.p[j0:j1], self.c[j0:j1] = p, sgn * t_excl * rinv
        self.sg[j0:j1], self.dn[j0:j1] = sgn * p * rinv, np.abs(dg) * t_incl * rinv
        self.t = t_incl[-1]
        self.S *= np.prod(sgn)
        buf = payload.numpy()
        m = j1 - j0
        buf[:] = 0
        buf[0:m], buf[RC_BLOCK:RC_BLOCK + m] = self.p[j0:j1], self.c[j0:j1]
        buf[2 * RC_BLOCK:2 * RC_BLOCK + m], buf[3 * RC_BLOCK:3 * RC_BLOCK + m] = self.sg[j0:j1], self.dn[j0:j1]
        buf[4 * RC_BLOCK], buf[4 * RC_BLOCK + 1] = self.t, self.S

    def panel_apply(self, J, is_owner, payload):
        j0, j1 = J * RC_BLOCK, min(self.n, (J + 1) * RC_BLOCK)
        m = j1 - j0
        if not is_owner:
            buf = payload.numpy()
            self.p[j0:j1], self.c[j0:j1] = buf[0:m], buf[RC_BLOCK:RC_BLOCK + m]
            self.sg[j0:j1], self.dn[j0:j1] = buf[2 * RC_BLOCK:2 * RC_BLOCK + m], buf[3 * RC_BLOCK:3 * RC_BLOCK + m]
            self.t, self.S = buf[4 * RC_BLOCK], buf[4 * RC_BLOCK + 1]
        sel = np.nonzero(self.rows >= j0)[0]
        if sel.size == 0:
            return
        R = self.rows[sel]
        w = self.w[R].copy()
        for k in range(j0, j1):
            below = R > k
            lo = self.A[sel, k].copy()
            self.A[sel[below], k] = self.c[k] * lo[below] + self.sg[k] * w[below]
            w[below] -= lo[below] * self.p[k]
            diag = R == k
            self.A[sel[diag], k] = self.dn[k]
        keep = R >= j1
        self.w[R[keep]] = w[keep]

    def status_tensor(self):
        return self.status

    # ---- downdate: forward substitution panel by panel, then the local reverse sweep
    def dd_panel_owner(self, J, payload):
        j0, j1 = J * RC_BLOCK, min(self.n, (J + 1) * RC_BLOCK)
        m = j1 - j0
        loc = np.searchsorted(self.rows, np.arange(j0, j1))
        LD = np.tril(self.A[loc, j0:j1])
        buf = payload.numpy()
        buf[:] = 0
        buf[0:m] = np.linalg.solve(LD, self.w[j0:j1])
        buf[3 * RC_BLOCK:3 * RC_BLOCK + m] = np.where(np.diag(LD) < 0, -1.0, 1.0)

    def dd_panel_apply(self, J, payload):
        j0, j1 = J * RC_BLOCK, min(self.n, (J + 1) * RC_BLOCK)
        m = j1 - j0
        buf = payload.numpy()
        self.p[j0:j1], self.dn[j0:j1] = buf[0:m], buf[3 * RC_BLOCK:3 * RC_BLOCK + m]
        sel = np.nonzero(self.rows >= j1)[0]
        if sel.size:
            self.w[self.rows[sel]] -= self.A[sel, j0:j1] @ self.p[j0:j1]

    def dd_finish(self):
        rho_sq = 1.0 - float(self.p @ self.p)
        if rho_sq <= 0:
            self.v_out[:] = self.p
            self.status[0] = 2
            return 2
        suffix = np.concatenate((np.cumsum((self.p * self.p)[::-1])[::-1], [0.0]))
        q = np.sqrt(rho_sq + suffix)  # q[k]^2 = rho^2 + sum_{j >= k} p_j^2
        c, s = q[1:] / q[:-1], self.p / q[:-1]
        for i, R in enumerate(self.rows):
            x = 0.0
            for k in range(R, -1, -1):
                lo = self.A[i, k]
                self.A[i, k] = self.dn[k] * (c[k] * lo - s[k] * x)
                x = c[k] * x + s[k] * lo
        self.v_out[:] = 0.25  # stands in for rotg's z (scratch)
        return 0

    # ---- wavefront over K successive updates (same slot layout as csrc/rowcyclic.cuh)
    PAY = 5 * RC_BLOCK + 8

    def _span(self, d):
        nJ = (self.n + RC_BLOCK - 1) // RC_BLOCK
        return max(0, d - self.K + 1), min(nJ - 1, d)

    def wave_begin(self, V, V_out, check_diag):
        self.K, self.V_out = int(V.shape[0]), V_out.numpy()
        nJ = (self.n + RC_BLOCK - 1) // RC_BLOCK
        self.max_tasks = max(1, -(-min(self.K, nJ) // self.world))
        self.W = np.zeros((self.K, self.n))
        self.W[:, self.rows] = V.numpy()[:, self.rows]
        self.tS = np.ones((self.K, 2))
        self.status[0] = 0
        if check_diag and np.any(self.A[np.arange(len(self.rows)), self.rows] == 0):
            self.status[0] = 1
        slots = self.max_tasks * self.PAY
        pay_all = torch.zeros(self.world * slots, dtype=torch.float64)
        pay_mine = pay_all[:slots] if self.world == 1 else torch.zeros(slots, dtype=torch.float64)
        return pay_mine, pay_all

    def wave_owner(self, d, pay_mine):
        Jlo, Jhi = self._span(d)
        Jfirst = Jlo + (self.rank - Jlo) % self.world
        buf = pay_mine.numpy()
        for i, J in enumerate(range(Jfirst, Jhi + 1, self.world)):
            u = d - J
            j0, j1 = J * RC_BLOCK, min(self.n, (J + 1) * RC_BLOCK)
            m = j1 - j0
            loc = np.searchsorted(self.rows, np.arange(j0, j1))
            LD = np.tril(self.A[loc, j0:j1])
            p = np.linalg.solve(LD, self.W[u, j0:j1])
            dg = np.diag(LD)
            t_in, S_in = self.tS[u]
            t_incl = t_in + np.cumsum(p * p)
            t_excl = np.concatenate(([t_in], t_incl[:-1]))
            sgn = np.where(dg < 0, -1.0, 1.0)
            rinv = 1.0 / np.sqrt(t_excl * t_incl)
            pay = buf[i * self.PAY:(i + 1) * self.PAY]
            pay[:] = 0
            pay[0:m] = p
            pay[RC_BLOCK:RC_BLOCK + m] = sgn * t_excl * rinv
            pay[2 * RC_BLOCK:2 * RC_BLOCK + m] = sgn * p * rinv
            pay[3 * RC_BLOCK:3 * RC_BLOCK + m] = np.abs(dg) * t_incl * rinv
            pay[4 * RC_BLOCK:4 * RC_BLOCK + m] = 0.5  # stands in for rotg's z (scratch)
            pay[5 * RC_BLOCK], pay[5 * RC_BLOCK + 1] = t_incl[-1], S_in * np.prod(sgn)

    def wave_apply(self, d, pay_all):
        Jlo, Jhi = self._span(d)
        buf = pay_all.numpy()
        for J in range(Jlo, Jhi + 1):
            u, t = d - J, J - Jlo
            slot = (J % self.world) * self.max_tasks + t // self.world
            pay = buf[slot * self.PAY:(slot + 1) * self.PAY]
            j0, j1 = J * RC_BLOCK, min(self.n, (J + 1) * RC_BLOCK)
            m = j1 - j0
            self.tS[u] = pay[5 * RC_BLOCK], pay[5 * RC_BLOCK + 1]
            self.V_out[u, j0:j1] = pay[4 * RC_BLOCK:4 * RC_BLOCK + m]
            sel = np.nonzero(self.rows >= j0)[0]
            if sel.size == 0:
                continue
            R = self.rows[sel]
            w = self.W[u, R].copy()
            for k in range(j0, j1):
                a = k - j0
                below = R > k
                lo = self.A[sel, k].copy()
                self.A[sel[below], k] = pay[RC_BLOCK + a] * lo[below] + pay[2 * RC_BLOCK + a] * w[below]
                w[below] -= lo[below] * pay[a]
                self.A[sel[R == k], k] = pay[3 * RC_BLOCK + a]
            keep = R >= j1
            self.W[u, R[keep]] = w[keep]
