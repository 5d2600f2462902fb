"""Lane-level emulation of k_gs_strip (saras_b200/csrc/gs_strip.cuh) on the CPU.

The kernel's schedule -- position / ring-slot arithmetic, prologue, prefetch distances, the three warp-uniform waits, the
comm warp, ticket order -- is restated here statement by statement and EXECUTED on small grids with real fp64 arithmetic:
warps are coroutines picked in random order (so every legal interleaving of tiles can occur, including slow comm warps),
cp.async copies and prefetching register loads are performed at a random moment between their issue and the point the
program order forces them to be complete (so a read that is too early OR too late shows up), and the result is compared
BIT FOR BIT with a plain sequential lexicographic Gauss-Seidel sweep of the same arithmetic (poisson_d3.cc:137-150).
A wrong slot, a missing dependency or a deadlock cannot pass.  Python floats are IEEE doubles and nothing is contracted,
so the comparison is exact.

    python scripts/sim/gs_strip_emul.py            # a set of shapes x geometries x seeds
"""
import random
import sys

import numpy as np

ZD = 2


class Grid:
    def __init__(self, nx, ny, nz, seed, zuni=True):
        rng = np.random.default_rng(seed)
        self.nx, self.ny, self.nz = nx, ny, nz
        self.NQ = (nz + 3) // 4
        self.kp = 4 * self.NQ + 8                     # z extent incl. the k = -1 pad at index 0 .. and slack above
        self.P = rng.uniform(-1, 1, (nx + 2, ny + 2, self.kp))
        self.R = rng.uniform(-1, 1, (nx + 2, ny + 2, self.kp))
        self.xix2 = rng.uniform(0.5, 1.5, nx + 2); self.xixx = rng.uniform(-0.2, 0.2, nx + 2)
        self.ety2 = rng.uniform(0.5, 1.5, ny + 2); self.etyy = rng.uniform(-0.2, 0.2, ny + 2)
        if zuni:
            self.ztz2 = np.full(self.kp, 1.25); self.ztzz = np.zeros(self.kp)
        else:
            self.ztz2 = rng.uniform(0.5, 1.5, self.kp); self.ztzz = rng.uniform(-0.2, 0.2, self.kp)
        self.zuni = zuni
        self.hx, self.hy, self.hz = 0.5, 0.25, 0.125
        self.hyhz, self.hzhx, self.hxhy = 0.03, 0.06, 0.12
        self.two_h = 2.0 * 0.0075

    # element (i, j, k) with i, j, k in -1 .. n
    def at(self, A, i, j, k):
        return A[i + 1, j + 1, k + 1]


def relax(g, pE, pW, pN, pS, pT, pB, rhs, i, j, k):
    cx2 = g.hyhz * g.xix2[i + 1]; cx1 = g.hyhz * g.xixx[i + 1]
    cy2 = g.hzhx * g.ety2[j + 1]; cy1 = g.hzhx * g.etyy[j + 1]
    cz2 = g.hxhy * g.ztz2[k + 1]; cz1 = g.hxhy * g.ztzz[k + 1]
    u = cx2 * (pE + pW) * 2.0
    u = u + cx1 * (pE - pW) * g.hx
    u = u + cy2 * (pN + pS) * 2.0
    u = u + cy1 * (pN - pS) * g.hy
    u = u + cz2 * (pT + pB) * 2.0
    u = u + cz1 * (pT - pB) * g.hz
    u = u - g.two_h * rhs
    return u / (4.0 * ((cx2 + cy2) + cz2))


def sequential_gs(g):
    P = g.P.copy()
    for i in range(g.nx):
        for j in range(g.ny):
            for k in range(g.nz):
                P[i + 1, j + 1, k + 1] = relax(g, P[i + 2, j + 1, k + 1], P[i, j + 1, k + 1], P[i + 1, j + 2, k + 1], P[i + 1, j, k + 1],
                                               P[i + 1, j + 1, k + 2], P[i + 1, j + 1, k], g.R[i + 1, j + 1, k + 1], i, j, k)
    return P


class Plan:
    def __init__(self, g, nq, DP, WC, LB=32):
        self.LB = LB                                  # lanes per strip (32 on the GPU; smaller here to keep grids tiny)
        self.NJ = (g.ny + LB - 1) // LB
        self.nq, self.DP, self.NQ = nq, DP, g.NQ
        self.NZW = (g.NQ + nq - 1) // nq
        self.NG = (self.NZW + WC - 1) // WC
        self.WC = (self.NZW + self.NG - 1) // self.NG
        self.RW, self.RRn = 2 * nq + DP + 1, DP + 1
        ts = sorted(((LB * J + nq * self.WC * G, G, J) for G in range(self.NG) for J in range(self.NJ)))
        self.tasks = [(J, G) for _, G, J in ts]
        self.sync = [0] * (1 + self.NJ * self.NZW)


class Pending:
    """a cp.async / prefetching load: performed at a random moment between issue and forced completion"""
    def __init__(self, fn, rng):
        self.fn, self.done = fn, False
        if rng.random() < 0.4:
            self.run()

    def run(self):
        if not self.done:
            self.fn(); self.done = True


def compute_warp(g, pl, cta, a, rng):
    LB = pl.LB
    J, Gz = cta["J"], cta["G"]
    nx, nq, NJ, WC, DP, RW, RRn = g.nx, pl.nq, pl.NJ, pl.WC, pl.DP, pl.RW, pl.RRn
    nb = min(LB, g.ny - LB * J)
    zw = Gz * WC + a
    if zw >= pl.NZW:
        return
    q0 = zw * nq
    nqw = min(nq, pl.NQ - q0)
    Pend = nx * nq                                    # every slab walks nq positions per plane; quads q >= nqw of the top slab are dead
    has_below_cta = a > 0
    has_below_far = (a == 0 and Gz > 0)
    has_above_cta = (a + 1 < WC) and (zw + 1 < pl.NZW)
    has_y = J > 0
    win = [[None] * (LB + 2) for _ in range(RW)]      # win[slot][cell] = quad; cell = lane + 1, cells 0 and nb + 1 hold the y halos
    rr = [[None] * LB for _ in range(RRn)]
    SD = (nq - 1 + DP) // nq + 1                      # side rings: planes per lane
    XW = 2
    while XW < (LB - 1 + DP) // nq + 2:
        XW *= 2
    xtab = [None] * XW                                # x coefficients of plane (i mod XW), requested by lane 0
    rdr = [[None] * LB for _ in range(SD)]            # (i, j) tag standing for the reciprocal-diagonal table entry
    ptr = [[None] * LB for _ in range(SD)]            # OLD first point above the slab
    pbr = [[None] * LB for _ in range(SD)]            # -z value from global memory (bottom slab: pad; first warp of a group: other CTA)
    zb_mine = cta["zb"][a]; zb_above = cta["zb"][a + 1] if a + 1 <= WC else None
    nvalid_last = g.nz - 4 * (pl.NQ - 1)
    t0 = -(2 * nq + DP)
    lanes = []
    for b in range(LB):
        col_ok = b < nb
        j = LB * J + (b if col_ok else nb - 1)
        pos = t0 - b
        m = -pos
        i = -((m + nq - 1) // nq); q = pos - i * nq
        pr = pos + DP; mr = -pr
        ir = -((mr + nq - 1) // nq); qr = pr - ir * nq
        lanes.append(dict(b=b, col_ok=col_ok, j=j, pos=pos, i=i, q=q, ir=ir, qr=qr, sw=pos % RW, sr=pos % RRn,
                          cx2=0.0, cx1=0.0, s_xy=0.0, cx2n=None, cx1n=None, pBn=None, pTopn=None, pTop=0.0, pB=0.0,
                          cur=[0.0] * 4, prev_out=[0.0] * 4, groups=[]))
    seenY = seenZ = seenA = 0
    t_end = Pend + nb - 1

    def gquad(A, i, j, qq):
        k0 = 4 * qq
        return [float(A[i + 1, j + 1, k0 + 1 + e]) for e in range(4)]

    for t in range(t0, t_end):
        # ---- (A)
        pym = t + DP
        needY = pym + LB if (has_y and 0 <= pym < Pend) else 0
        while seenY < needY:
            seenY = cta["seenY"][a]
            if seenY < needY:
                yield "wait"
        final_all = Pend + nb - 1
        needZ = min(t + nq + (0 if has_below_cta else DP), final_all) if (has_below_cta or has_below_far) else 0
        needA = max(t + 2 - (ZD + 1) * nq, 0) if has_above_cta else 0
        for ln in lanes:
            wrap1 = ln["q"] + 1 == nq
            ln["q1"] = 0 if wrap1 else ln["q"] + 1
            ln["i1"] = ln["i"] + 1 if wrap1 else ln["i"]
            ln["next_live"] = ln["col_ok"] and 0 <= ln["i1"] < nx and ln["q1"] < nqw
            ln["live"] = ln["col_ok"] and 0 <= ln["i"] < nx and ln["q"] < nqw
        while seenZ < needZ:
            seenZ = cta["done"][a - 1] if has_below_cta else cta["seenZ0"]
            if seenZ < needZ:
                yield "wait"
        while seenA < needA:
            seenA = cta["done"][a + 1]
            if seenA < needA:
                yield "wait"
        # ---- (B) prefetch
        for ln in lanes:
            grp = []
            b, j, ir, qr = ln["b"], ln["j"], ln["ir"], ln["qr"]
            if ln["col_ok"] and qr < nqw and -1 <= ir + 1 <= nx:
                s = (ln["sw"] + nq + DP) % RW
                grp.append(Pending((lambda s=s, b=b, i=ir + 1, j=j, qq=q0 + qr: win[s].__setitem__(b + 1, gquad(g.P, i, j, qq))), rng))
            if ln["col_ok"] and qr < nqw and 0 <= ir < nx:
                s = (ln["sr"] + DP) % RRn
                grp.append(Pending((lambda s=s, b=b, i=ir, j=j, qq=q0 + qr: rr[s].__setitem__(b, gquad(g.R, i, j, qq))), rng))
                sh = (ln["sw"] + DP) % RW                # slot of position pos + DP
                if b == 0:
                    grp.append(Pending((lambda sh=sh, i=ir, j=j, qq=q0 + qr: win[sh].__setitem__(0, gquad(g.P, i, j - 1, qq))), rng))
                if b == nb - 1:
                    grp.append(Pending((lambda sh=sh, b=b, i=ir, j=j, qq=q0 + qr: win[sh].__setitem__(b + 2, gquad(g.P, i, j + 1, qq))), rng))
            if ln["col_ok"] and qr < nqw and 0 <= ir < nx:      # side rings of position pos + DP
                sd = ir % SD
                if qr == 0:
                    grp.append(Pending((lambda sd=sd, b=b, i=ir, j=j: rdr[sd].__setitem__(b, (i, j))), rng))
                    if not has_below_cta:
                        grp.append(Pending((lambda sd=sd, b=b, i=ir, j=j: pbr[sd].__setitem__(b, float(g.at(g.P, i, j, 4 * q0 - 1)))), rng))
                    if b == 0:
                        grp.append(Pending((lambda i=ir: xtab.__setitem__(i % XW, (i, float(g.xix2[i + 1]), float(g.xixx[i + 1])))), rng))
                if qr == nqw - 1:
                    grp.append(Pending((lambda sd=sd, b=b, i=ir, j=j: ptr[sd].__setitem__(b, float(g.at(g.P, i, j, 4 * (q0 + nqw))))), rng))
            ln["groups"].append(grp)
        # ---- wait_group(DP): groups of steps <= t - DP complete
        for ln in lanes:
            while len(ln["groups"]) > DP:
                for pnd in ln["groups"].pop(0):
                    pnd.run()
        # ---- (D)
        nxt = [win[(ln["sw"] + 1) % RW][ln["b"] + 1] for ln in lanes]
        for ln in lanes:
            b = ln["b"]
            if not ln["live"]:
                continue
            i, q, j = ln["i"], ln["q"], ln["j"]
            # ---- (C) entering a plane / at the top quad: side rings (filled DP steps ago) and the -z hand-off ring
            if q == 0:
                xt = xtab[i % XW]
                assert xt is not None and xt[0] == i, "x-coefficient ring holds the wrong plane"
                ln["cx2"], ln["cx1"] = g.hyhz * xt[1], g.hyhz * xt[2]
                ln["s_xy"] = ln["cx2"] + g.hzhx * g.ety2[j + 1]
                assert rdr[i % SD][b] == (i, j), "reciprocal-diagonal ring holds the wrong column"
                ln["pB"] = zb_mine[i % ZD][b] if has_below_cta else pbr[i % SD][b]
                assert ln["pB"] is not None
            if q == nqw - 1:
                ln["pTop"] = ptr[i % SD][b]
                assert ln["pTop"] is not None
            wp = win[(ln["sw"] + nq) % RW][b + 1]
            wm = win[(ln["sw"] + RW - nq) % RW][b + 1]
            rq = rr[ln["sr"]][b]
            ymq = win[ln["sw"]][b]                     # left cell of the own slot row: lane b-1's NEW quad / the -y halo
            ypq = win[ln["sw"]][b + 2]                 # right cell: lane b+1's OLD quad / the +y halo
            assert wp is not None and wm is not None and rq is not None and ymq is not None and ypq is not None and (q == nqw - 1 or nxt[b] is not None), "unfilled ring slot read"
            nvalid = nvalid_last if q0 + q == pl.NQ - 1 else 4
            cy2 = g.hzhx * g.ety2[j + 1]; cy1 = g.hzhx * g.etyy[j + 1]
            cx2, cx1 = ln["cx2"], ln["cx1"]
            cur = win[ln["sw"]][b + 1]                   # the quad's own OLD values, still in its window cell
            assert cur is not None
            pT3 = ln["pTop"] if q == nqw - 1 else nxt[b][0]
            o = [0.0] * 4
            pb = ln["pB"]
            for e in range(4):
                k = 4 * (q0 + q) + e
                cz2 = g.hxhy * g.ztz2[k + 1] if k < g.nz else 0.0
                cz1 = g.hxhy * g.ztzz[k + 1] if k < g.nz else 0.0
                u = cx2 * (wp[e] + wm[e]) * 2.0
                u = u + cx1 * (wp[e] - wm[e]) * g.hx
                u = u + cy2 * (ypq[e] + ymq[e]) * 2.0
                u = u + cy1 * (ypq[e] - ymq[e]) * g.hy
                pT = cur[e + 1] if e < 3 else pT3
                u = u + cz2 * (pT + pb) * 2.0
                u = u + cz1 * (pT - pb) * g.hz
                u = u - g.two_h * rq[e]
                den = 4.0 * (ln["s_xy"] + cz2)
                r = u / den if den != 0.0 else 0.0
                if e < nvalid:
                    pb = r
                o[e] = r
            ln["pB"] = pb
            for e in range(nvalid):
                g.P[i + 1, j + 1, 4 * (q0 + q) + e + 1] = o[e]
            win[ln["sw"]][b + 1] = list(o)
            if has_above_cta and q == nqw - 1:
                zb_above[i % ZD][b] = o[3]
        # ---- (E)
        for ln in lanes:
            ln["pos"] += 1
            ln["q"], ln["i"] = ln["q1"], ln["i1"]
            ln["qr"] += 1
            if ln["qr"] == nq:
                ln["qr"] = 0; ln["ir"] += 1
            ln["sw"] = (ln["sw"] + 1) % RW
            ln["sr"] = (ln["sr"] + 1) % RRn
        if t >= 0:
            cta["done"][a] = t + 1
        yield "step"


def comm_warp(g, pl, cta):
    J, Gz, WC, nq, NJ, nx = cta["J"], cta["G"], pl.WC, pl.nq, pl.NJ, g.nx
    nb = min(pl.LB, g.ny - pl.LB * J)
    pub = [0] * WC
    while True:
        fin = True
        polled = []
        for w in range(WC):
            zw = Gz * WC + w
            if zw >= pl.NZW:
                polled.append(None); continue
            polled.append((cta["done"][w], pl.sync[1 + zw * NJ + J - 1] if J > 0 else 0))
        z = pl.sync[1 + (Gz * WC - 1) * NJ + J] if Gz > 0 else 0
        yield "comm"                                   # the fence: anything may happen in between
        for w in range(WC):
            if polled[w] is None:
                continue
            zw = Gz * WC + w
            d, y = polled[w]
            if d > pub[w]:
                pl.sync[1 + zw * NJ + J] = d; pub[w] = d
            if y > cta["seenY"][w]:
                cta["seenY"][w] = y
            if pub[w] < nx * nq + nb - 1:
                fin = False
        if z > cta["seenZ0"]:
            cta["seenZ0"] = z
        if fin:
            return
        yield "comm"


def run_case(nx, ny, nz, nq, DP, WC, LB, seed, slots=10 ** 9, zuni=True):
    rng = random.Random(seed)
    g = Grid(nx, ny, nz, seed, zuni)
    ref = sequential_gs(g)
    pl = Plan(g, nq, DP, WC, LB)
    ticket = 0
    active = []

    def admit():
        nonlocal ticket
        while ticket < len(pl.tasks) and len({id(c) for c, _ in active}) < slots:
            J, G = pl.tasks[ticket]; ticket += 1
            cta = dict(J=J, G=G, done=[0] * (pl.WC + 2), seenY=[0] * (pl.WC + 1), seenZ0=0,
                       zb=[[[None] * LB for _ in range(ZD)] for _ in range(pl.WC + 2)])
            for a in range(pl.WC):
                active.append((cta, compute_warp(g, pl, cta, a, rng)))
            active.append((cta, comm_warp(g, pl, cta)))

    admit()
    idle = 0
    while active:
        k = rng.randrange(len(active))
        cta, gen = active[k]
        try:
            r = next(gen)
            idle = idle + 1 if r == "wait" or r == "comm" else 0
        except StopIteration:
            active.pop(k)
            admit()
            idle = 0
        assert idle < 200000, "deadlock: no compute warp advanced in 200000 scheduler picks"
    got = g.P
    core = (slice(1, nx + 1), slice(1, ny + 1), slice(1, nz + 1))
    assert np.array_equal(got[core], ref[core]), "emulated strip sweep differs from sequential Gauss-Seidel: max |d| = %g" % np.max(np.abs(got[core] - ref[core]))
    untouched = np.ones(got.shape, bool); untouched[core] = False
    assert np.array_equal(got[untouched], ref[untouched]), "pads were written"
    return True


CASES = [
    # nx, ny, nz, nq, DP, WC, LB
    (4, 5, 9, 2, 2, 2, 4),        # 2 strips (4 + 1 lanes), 3 quads -> 2 slabs (2 + 1), one CTA per strip
    (5, 9, 17, 2, 2, 2, 4),       # 3 strips, 5 quads -> 3 slabs in 2 groups: cross-CTA -z
    (3, 8, 13, 1, 2, 3, 4),       # nq = 1: every quad is a plane start and a slab top; nz % 4 == 1
    (6, 7, 16, 3, 3, 2, 4),       # nz % 4 == 0: the +z pad is the first point of the quad above the column
    (4, 6, 33, 4, 2, 2, 3),       # taller slabs, lanes = 3
    (7, 4, 12, 2, 3, 1, 4),       # one warp per CTA: every -z hand-off crosses CTAs
    (3, 12, 11, 5, 2, 2, 4),      # a single slab taller than the column (nqw < nq)
]


def main():
    runs = 0
    for (nx, ny, nz, nq, DP, WC, LB) in CASES:
        for seed in range(3):
            for slots in (10 ** 9, 2):
                run_case(nx, ny, nz, nq, DP, WC, LB, seed, slots, zuni=(seed != 1))
                runs += 1
    print("gs_strip_emul: %d emulated sweeps bit-identical to sequential lexicographic Gauss-Seidel, no deadlock" % runs)


if __name__ == "__main__":
    sys.exit(main())
