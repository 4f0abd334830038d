"""D's old-time levels for the `backward` schemes (shipped 3dTube / HronTurek solids: d2dt2(D) backward, ddt(D) backward).

numerics/backwardD2dt2Scheme/backwardD2dt2Scheme.C:58-76 and [EXT] backwardDdtScheme::deltaT0_ decide from the STATE of the
field's old-time levels -- how many exist, which timeIndex each carries -- whether a use of the scheme sees the real
runTime.deltaT0() or GREAT (a first-order start-up).  In a foam-extend process that state lives in GeometricField (oldTime()
creates a level on first use, storeOldTimes() shifts the levels on the first non-const access of a new time step).  Here the
levels are device fields of the mesh handle (Dold, Doldold, Dooo, Doooo) and this class is the host-side bookkeeping that goes
with them: it creates / shifts them with gpupcg_mesh_copy_field exactly when GeometricField would and returns the effective
old time step of every use for gpupcg_eqn_controls (include/gpupcg.h)."""

GREAT = 1.0e15
LEVELS = ("Dold", "Doldold", "Dooo", "Doooo")


class OldTimes(object):
    def __init__(self, gm, firstTimeIndex=1):
        self.gm = gm
        self.ti = []                    # timeIndex of the existing levels, Dold first
        self.tiD = firstTimeIndex       # D.timeIndex()
        self.now = firstTimeIndex       # runTime.timeIndex()

    def advance(self):
        self.now += 1

    def touch(self):
        """First non-const access of D in a time step (the fvMatrix constructor): storeOldTimes() -- deepest level first, each
        takes the values and the timeIndex of the one above."""
        if self.ti and self.tiD != self.now:
            for j in range(len(self.ti) - 1, 0, -1):
                self.gm.copy_field(LEVELS[j], LEVELS[j - 1])
                self.ti[j] = self.ti[j - 1]
            self.gm.copy_field(LEVELS[0], "D")
            self.ti[0] = self.tiD
        self.tiD = self.now

    def ensure(self, k):
        """oldTime() down to level k: a missing level is created as a copy of the one above, with its timeIndex."""
        while len(self.ti) < k:
            j = len(self.ti)
            self.gm.copy_field(LEVELS[j], "D" if j == 0 else LEVELS[j - 1])
            self.ti.append(self.tiD if j == 0 else self.ti[j - 1])

    def n_old(self, j=0):
        return max(len(self.ti) - j, 0)

    def assembly(self, d2dt2, ddt, damping, deltaT0):
        """The uses of one D-equation assembly in the reference's order (unsTotalLagrangianStress.C:846-865)."""
        rule = lambda n: GREAT if n < 2 else deltaT0
        e = dict(eD2=deltaT0, eDdtD=deltaT0, eDdtD0=deltaT0, eDdtD00=deltaT0, eDamp=deltaT0)
        self.touch()
        if d2dt2 == "backward":
            self.ensure(3)                                              # deltaT0_(vf): vf.oldTime().oldTime().oldTime()
            e["eD2"] = GREAT if self.ti[1] == self.ti[2] else deltaT0
            e["eDdtD"] = rule(self.n_old(0))
            e["eDdtD0"] = rule(self.n_old(1))
            e["eDdtD00"] = rule(self.n_old(2))                          # fvcDdt(D.oldTime().oldTime()) then creates the 4th level
            self.ensure(4)
        elif d2dt2 == "Euler":
            self.ensure(2)
        if damping:
            if ddt == "backward":
                e["eDamp"] = rule(self.n_old(0))
                self.ensure(2)
            elif ddt == "Euler":
                self.ensure(1)
        return e

    def velocity(self, deltaT0):
        """U = fvc::ddt(D) after the outer loop (:973): residual() has used D.oldTime() by then."""
        self.ensure(1)
        eU = GREAT if self.n_old(0) < 2 else deltaT0
        self.ensure(2)
        return eU
