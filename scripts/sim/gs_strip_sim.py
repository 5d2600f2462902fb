"""Schedule simulator for the planned Gauss-Seidel sweep: no arithmetic, only WHO READS WHAT WHEN.

Every grid point carries a version (0 = value before the sweep, 1 = after).  Lexicographic Gauss-Seidel requires that the
update of (i,j,k) reads version 1 of (i-1,j,k), (i,j-1,k), (i,j,k-1) and version 0 of (i+1,j,k), (i,j+1,k), (i,j,k+1)
(pads are static).  The simulator executes the strip schedule -- warps as independent agents that advance one step at a
time whenever their wait conditions hold, picked in random order -- and checks every read against that rule, that every
point is written exactly once, and that the schedule never deadlocks, including with fewer resident CTA slots than CTAs.

Schedule under test
  strip     = (plane i, LB consecutive columns j of y-tile J, the quads of z-slab Z); a warp runs a strip in
              S = nq(Z) + LB - 1 steps, lane b handling quad t - b at step t (k recurrence in the lane, -y / +y
              neighbours from the adjacent lanes: produced one step earlier / not touched until one step later);
  CTA       = (J, Z), W warps; warp w runs the strips of planes w, w+W, w+2W, ...;
  -x quads  : ring written by the warp of plane i-1; the consumer needs the producer to have COMPLETED the same step
              index (so it runs >= 1 step behind), the producer may run at most `depth` steps ahead of the consumer.
              The W warps of a CTA form a CIRCULAR pipeline (warp 0 consumes warp W-1's planes one round later), and the
              lags around the circle add up to one strip length S, so the ring depth must be at least ceil(S / W): the
              simulator deadlocks with less.  depth = ceil(S / W) + R, R >= 1 being the elasticity;
  -y halo   : lane 0 reads column LB*J-1 of CTA (J-1, Z): needs that CTA's progress flag of plane i >= q + LB
              (flags are published every G steps and at the end of a strip);
  -z halo   : the strip may start when CTA (J, Z-1) has finished the strip of the same plane;
  residency : CTAs take tickets in the order of J + Z (ties: Z, J); at most SLOTS CTAs are resident; a CTA keeps its slot
              until it has run all its planes.

    python scripts/sim/gs_strip_sim.py            # sweeps a set of shapes and random schedules, prints a summary
"""
import itertools
import random
import sys


class Sim:
    def __init__(self, nx, ny, nz, LB, W, NZ, R, G, slots, seed):
        self.nx, self.ny, self.nz, self.LB, self.W, self.R, self.G = nx, ny, nz, LB, W, R, G
        self.NQ = (nz + 3) // 4
        self.NJ = (ny + LB - 1) // LB
        self.NZ = min(NZ, self.NQ)
        zq = self.NQ // self.NZ
        self.q0 = [z * zq for z in range(self.NZ)]
        self.nq = [(self.NQ - self.q0[z]) if z == self.NZ - 1 else zq for z in range(self.NZ)]
        self.ver = [[[0] * nz for _ in range(ny)] for _ in range(nx)]
        self.writes = 0
        self.rng = random.Random(seed)
        self.slots = slots
        # tickets: dependency-compatible order
        self.order = sorted(((J, Z) for J in range(self.NJ) for Z in range(self.NZ)), key=lambda c: (c[0] + c[1], c[1], c[0]))
        self.next_ticket = 0
        self.resident = {}          # (J,Z) -> cta state
        self.flag = {}              # (J,Z,i) -> published steps of the strip of plane i
        self.done_ctas = set()

    def steps(self, Z):
        return self.nq[Z] + self.LB - 1

    def admit(self):
        while len(self.resident) < self.slots and self.next_ticket < len(self.order):
            J, Z = self.order[self.next_ticket]
            self.next_ticket += 1
            # warp state: plane index (None when out of planes), step, completed absolute step count
            warps = []
            for w in range(self.W):
                warps.append(dict(w=w, plane=w if w < self.nx else None, t=0, abs=0))
            self.resident[(J, Z)] = dict(J=J, Z=Z, warps=warps)

    # absolute step index of (plane i, step t) in the timeline of the warp that owns plane i
    def abs_index(self, i, t, Z):
        return (i // self.W) * self.steps(Z) + t

    def can_step(self, cta, wp):
        J, Z, i, t = cta["J"], cta["Z"], wp["plane"], wp["t"]
        if i is None:
            return False
        S = self.steps(Z)
        if t == 0 and Z > 0 and self.flag.get((J, Z - 1, i), 0) < self.steps(Z - 1):
            return False                                    # -z halo: slab below finished for this plane
        if i > 0:                                           # -x ring: producer completed the same step
            prod = cta["warps"][(i - 1) % self.W]
            need = self.abs_index(i - 1, t, Z) + 1
            if prod["abs"] < need:
                return False
        if i + 1 < self.nx:                                 # back-pressure: at most R steps ahead of the consumer of my ring
            cons = cta["warps"][(i + 1) % self.W]
            mine = self.abs_index(i, t, Z)
            # consumer's completed step count measured on MY timeline: it consumes slot (plane i, step u) at its step u of plane i+1
            cplane, ct = cons["plane"], cons["t"]
            if cplane is None or cplane > i + 1:
                consumed = mine + 10 ** 9                    # consumer is past plane i+1 entirely
            elif cplane == i + 1:
                consumed = self.abs_index(i, ct, Z)         # it has consumed my steps < ct
            else:
                consumed = self.abs_index(i, 0, Z)          # it has not started plane i+1
            if mine - consumed >= (S + self.W - 1) // self.W + self.R:
                return False
        q = t                                               # lane 0's quad at this step
        if J > 0 and q < self.nq[Z]:                        # -y halo for lane 0
            if self.flag.get((J - 1, Z, i), 0) < min(q + self.LB, self.steps(Z)):
                return False
        return True

    def expect(self, i, j, k, v, what):
        if 0 <= i < self.nx and 0 <= j < self.ny and 0 <= k < self.nz:
            assert self.ver[i][j][k] == v, "%s of (%d,%d,%d) has version %d, expected %d" % (what, i, j, k, self.ver[i][j][k], v)

    def do_step(self, cta, wp):
        J, Z, i, t = cta["J"], cta["Z"], wp["plane"], wp["t"]
        S = self.steps(Z)
        for b in range(self.LB):
            j = J * self.LB + b
            q = t - b
            if j >= self.ny or q < 0 or q >= self.nq[Z]:
                continue
            for e in range(4):
                k = 4 * (self.q0[Z] + q) + e
                if k >= self.nz:
                    break
                self.expect(i - 1, j, k, 1, "-x neighbour")
                self.expect(i, j - 1, k, 1, "-y neighbour")
                self.expect(i, j, k - 1, 1, "-z neighbour")
                self.expect(i + 1, j, k, 0, "+x neighbour")
                self.expect(i, j + 1, k, 0, "+y neighbour")
                self.expect(i, j, k + 1, 0, "+z neighbour")
                assert self.ver[i][j][k] == 0, "(%d,%d,%d) written twice" % (i, j, k)
                self.ver[i][j][k] = 1
                self.writes += 1
        wp["t"] += 1
        wp["abs"] += 1
        if wp["t"] % self.G == 0 or wp["t"] == S:
            self.flag[(J, Z, i)] = wp["t"]                  # publication (granularity G)
        if wp["t"] == S:
            nxt = i + self.W
            wp["plane"] = nxt if nxt < self.nx else None
            wp["t"] = 0

    def run(self):
        self.admit()
        total = self.nx * self.ny * self.nz
        guard = 0
        while self.writes < total or self.resident:
            ready = [(c, w) for c in self.resident.values() for w in c["warps"] if self.can_step(c, w)]
            if not ready:
                finished = [key for key, c in self.resident.items() if all(w["plane"] is None for w in c["warps"])]
                if finished:
                    for key in finished:
                        del self.resident[key]
                    self.admit()
                    continue
                raise AssertionError("deadlock: %d of %d points written, %d CTAs resident, ticket %d/%d" % (
                    self.writes, total, len(self.resident), self.next_ticket, len(self.order)))
            c, w = self.rng.choice(ready)
            self.do_step(c, w)
            for key in [key for key, cc in self.resident.items() if all(x["plane"] is None for x in cc["warps"])]:
                del self.resident[key]
            self.admit()
            guard += 1
            assert guard < 10 ** 8
        assert self.writes == total
        return guard


def main():
    shapes = [(5, 9, 9), (7, 10, 17), (4, 3, 5), (9, 17, 33), (3, 5, 9), (6, 8, 12)]
    runs = 0
    for (nx, ny, nz), LB, W, NZ, R, G in itertools.product(shapes, (2, 4), (1, 2, 3), (1, 2, 3), (1, 3), (1, 3)):
        for slots in (1, 2, 1000):
            for seed in range(3):
                Sim(nx, ny, nz, LB, W, NZ, R, G, slots, seed).run()
                runs += 1
    print("gs_strip_sim: %d schedules executed, every read saw the version lexicographic Gauss-Seidel requires, no deadlock" % runs)


if __name__ == "__main__":
    sys.exit(main())
