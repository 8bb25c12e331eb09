"""Oracle (test infrastructure): OptimalControl / ControlBasis / seeds, NumPy restatement.

Follows ``src/OptimalControl.cpp`` (cache state machine and formulas, cited per method),
``src/ControlBasis.cpp``, ``include/ControlBasisFactory.hpp`` and ``include/SeedGenerator.hpp``.
"""
from __future__ import annotations

import math
import numpy as np

from .bh_mps import MPS, BHStepper, TruncArgs, overlap, overlap_K, apply_K

PI_REF = 3.14159265          # include/ControlBasisFactory.hpp:10 (truncated pi is behaviour)


# ----------------------------------------------------------------------------------------------
# include/SeedGenerator.hpp
# ----------------------------------------------------------------------------------------------
def linspace(a, b, n):
    """SeedGenerator::linspace (:26-37): accumulating loop with the 1e-7 guard."""
    out = []
    step = (b - a) / (n - 1)
    while a <= b + 1e-7:
        out.append(a)
        a += step
    return out


def generate_range(a, b, c):
    out = []
    while a <= c + 1e-7:
        out.append(a)
        a += b
    return out


def sigmoid(x, k, offset):
    return [1.0 / (1.0 + math.exp(-k * (xv - offset))) for xv in x]


def linsigmoid_seed(u_start, u_end, length, a, c, d):
    """SeedGenerator::linsigmoidSeed (:66-95) with the three random coefficients passed in
    (the reference draws them from libc rand(); a in [0.01,0.15], c in [0.06,0.18], d in [60,80])."""
    x = linspace(0, 100, length)
    b = u_end - u_start - a * x[-1]
    S1 = sigmoid(x, 0.7, 5)
    S2 = sigmoid(x, -0.9, 100 - 7)
    for i in range(len(S1) // 2, len(S1)):
        S1[i] = S2[i]
    S1[0] = 0
    S1[-1] = 0
    out = []
    for i, fx in enumerate(x):
        out.append(S1[i] * (a * fx + b / (1 + math.exp(-c * (fx - d))) + u_start)
                   + (1 - S1[i]) * ((u_end - u_start) / (1 + math.exp(-0.2 * (fx - 40))) + u_start))
    return out


def adiabatic_seed(u_start, u_end, length):
    xs_ = linspace(0, 100, length)
    p, k, xs, a = 3.5, 1.0 / 3.0, 40, 0.01
    out = []
    for x in xs_:
        if x < xs:
            out.append((p - u_start - a * xs) / (1 + math.exp(-k * (x - xs / 2.0))) + u_start + a * x)
        else:
            out.append(math.exp(math.log(u_end - p + 1) / (100 - xs) * (x - xs)) + p - 1)
    return out


# ----------------------------------------------------------------------------------------------
# src/ControlBasis.cpp
# ----------------------------------------------------------------------------------------------
class ControlBasis:
    def __init__(self, u0=None, S=None, f=None):
        if u0 is None:
            self.N = self.M = 0
            return
        self.u0 = list(map(float, u0))
        self.S = list(map(float, S))
        self.f = [list(map(float, row)) for row in f]
        self.N = len(self.u0)
        self.M = len(self.f[0])
        self.jac = [[self.f[i][n] * self.S[i] for n in range(self.M)] for i in range(self.N)]   # :14-24
        self.ucurrent = list(self.u0)                                                            # :27
        self.vmat = [[self.jac[j][i] for j in range(self.N)] for i in range(self.M)]            # :30-38

    def getM(self):
        return self.M

    def getN(self):
        return self.N

    def convertControl(self, control, new_control=True):       # :49-67
        if new_control:
            assert len(control) == self.M
            u = list(self.u0)
            for i in range(self.N):
                acc = 0.0
                for n in range(self.M):
                    acc += self.f[i][n] * control[n]
                u[i] += self.S[i] * acc
            self.ucurrent = u
        return list(self.ucurrent)

    def convertGradient(self, gradu):                          # :70-89
        assert len(gradu) == self.N
        out = []
        for n in range(self.M):
            gn = 0.0
            for i in range(self.N):
                gn += self.S[i] * gradu[i] * self.f[i][n]
            out.append(gn)
        return out

    def convertHessian(self, Hu):                              # :92-119
        Hu = np.asarray(Hu, dtype=float)
        V = np.asarray(self.vmat, dtype=float)
        Hc = np.zeros((self.M, self.M))
        for i in range(self.M):
            for j in range(i, self.M):
                Hvj = Hu @ V[j]
                Hc[i, j] = float(V[i] @ Hvj)
                Hc[j, i] = Hc[i, j]
        return Hc

    def getControlJacobian(self):                              # :122
        return [list(r) for r in self.jac]


def build_chopped_sine_basis(u0, tstep, T, M):
    """ControlBasisFactory::buildChoppedSineBasis (include/ControlBasisFactory.hpp:25-52)."""
    N = len(u0)
    x = linspace(0, 100, N)
    S = sigmoid(x, 8.0, 1.1)
    S2 = sigmoid(x, -8.0, 100 - 1.1)
    for i in range(N // 2, N):
        S[i] = S2[i]
    S[0] = 0
    S[N - 1] = 0
    f = [[math.sin((n + 1) * PI_REF * tstep * i / T) for n in range(M)] for i in range(N)]
    return ControlBasis(u0, S, f)


# ----------------------------------------------------------------------------------------------
# src/OptimalControl.cpp
# ----------------------------------------------------------------------------------------------
class OptimalControl:
    """OptimalControl<BH_tDMRG>.  ``basis=None`` -> GRAPE constructor (:7-28) with ``N`` time
    points, otherwise the GROUP constructor (:31-52)."""

    def __init__(self, psi_target: MPS, psi_init: MPS, stepper: BHStepper, N=None, basis=None,
                 gamma=0.0, BFGS=False):
        self.psi_target = psi_target.copy()
        self.psi_init = psi_init.copy()
        self.stepper = stepper
        self.tstep = stepper.get_tstep()
        self.gamma = gamma
        self.BFGS = BFGS
        self.calculatedXi = False
        self.threadCount = 1
        if basis is None:
            self.basis = ControlBasis()
            self.GRAPE = True
            self.N = int(N)
            self.M = 0
        else:
            self.basis = basis
            self.GRAPE = False
            self.N = basis.getN()
            self.M = basis.getM()
        self.psi_t = [None] * self.N
        self.divT = [0j] * self.N
        self.xi_t = [] if BFGS else [None] * self.N
        self.xiHlist = [] if BFGS else [None] * self.N
        self.n_steps = 0      # bookkeeping for tests/bench: Trotter steps executed

    # -- setters :55-85
    def setThreadCount(self, n):
        if n < 1:
            raise ValueError("Mininum threadCount is 1.")
        self.threadCount = n

    def setGRAPE(self, use):
        self.GRAPE = use
        self.calculatedXi = False

    def setBFGS(self, use):
        self.BFGS = use
        self.calculatedXi = False
        if use:
            self.xi_t, self.xiHlist = [], []
        else:
            self.xi_t, self.xiHlist = [None] * self.N, [None] * self.N

    def useBFGS(self):
        return self.BFGS

    def setGamma(self, g):
        self.gamma = g

    def getM(self):
        return self.M

    def getN(self):
        return self.N

    def getPsit(self):
        return list(self.psi_t)

    def getControl(self, control):
        return list(control) if self.GRAPE else self.basis.convertControl(control)

    def getTimeAxis(self):                                   # :188-201
        out, t = [], 0.0
        while abs(t - self.N * self.tstep) > 1e-2 * self.tstep:
            out.append(t)
            t += self.tstep
        return out

    # -- regularisation :89-143
    def calcRegularization(self, u):
        tmp = 0.0
        for i in range(self.N - 1):
            d = u[i + 1] - u[i]
            tmp += d * d / self.tstep
        return self.gamma / 2.0 * tmp

    def calcRegularizationGrad(self, u):
        N, g, t = self.N, self.gamma, self.tstep
        out = [-g * (-5.0 * u[1] + 4.0 * u[2] - u[3] + 2.0 * u[0]) / t]
        for i in range(1, N - 1):
            out.append(-g * (u[i + 1] + u[i - 1] - 2.0 * u[i]) / t)
        out.append(-g * (-5.0 * u[N - 2] + 4.0 * u[N - 3] - u[N - 4] + 2.0 * u[N - 1]) / t)
        return out

    def calcRegularizationHessian(self, u):
        N = self.N
        H = np.zeros((N, N))
        got = self.gamma / self.tstep
        for i in range(1, N - 1):
            H[i, i - 1] = -got
            H[i, i + 1] = -got
            H[i, i] = 2.0 * got
        H[1, 0] = 0
        H[N - 2, N - 1] = 0
        return H

    # -- sweeps :376-438
    def calcPsi(self, u):
        psi0 = self.psi_init.copy()
        self.psi_t[0] = psi0.copy()
        for i in range(self.N - 1):
            self.stepper.step(psi0, u[i], u[i + 1], True)
            self.n_steps += 1
            self.psi_t[i + 1] = psi0.copy()
        self.calculatedXi = False

    def calcXi(self, u):
        xiT = self.psi_target.copy()
        self.xi_t[self.N - 1] = xiT.copy()
        for i in range(self.N - 1, 0, -1):
            self.stepper.step(xiT, u[i], u[i - 1], False)
            self.n_steps += 1
            self.xi_t[i - 1] = xiT.copy()
        self.calculatedXi = True

    def calcDivT(self, u):
        assert self.calculatedXi
        for i in range(self.N):
            self.divT[i] = overlap_K(self.xi_t[i], self.psi_t[i])

    def calcPsiXiDivT(self, u):
        self.calcPsi(u)
        self.calcXi(u)
        self.calcDivT(u)

    # -- cost :441-453
    def calcCost(self, u, new_control=True):
        if new_control:
            self.calculatedXi = False
            self.calcPsi(u)
        ov = overlap(self.psi_target, self.psi_t[-1])
        return 0.5 * (1.0 - (ov.real ** 2 + ov.imag ** 2)) + self.calcRegularization(u)

    # -- gradient :205-249, :457-467
    def calcFidelityGrad(self, u, new_control=True):
        if new_control:
            self.calculatedXi = False
            if self.BFGS:
                self.calcPsi(u)
            else:
                self.calcPsiXiDivT(u)
        if self.BFGS:
            xi = self.psi_target.copy()
            self.divT[self.N - 1] = overlap_K(xi, self.psi_t[-1])
            for i in range(self.N - 1, 0, -1):
                self.stepper.step(xi, u[i], u[i - 1], False)
                self.n_steps += 1
                self.divT[i - 1] = overlap_K(xi, self.psi_t[i - 1])
        else:
            if not self.calculatedXi:
                self.calcXi(u)
                self.calcDivT(u)
        of = overlap(self.psi_t[-1], self.psi_target)
        return [self.tstep * (self.divT[i] * of * 1j).real for i in range(self.N)]

    def calcAnalyticGradient(self, u, new_control=True):
        fg = self.calcFidelityGrad(u, new_control)
        rg = self.calcRegularizationGrad(u)
        return [a + b for a, b in zip(fg, rg)]

    # -- Hessian :252-372
    def calcHessianRow(self, row, u, of, H):
        args = self.stepper.args
        psiH = apply_K(self.psi_t[row], args)
        normiH = psiH.norm()
        ts2 = self.tstep * self.tstep
        val1 = (of * overlap(self.xiHlist[row], psiH)).real
        val2 = -(self.divT[row] * np.conj(self.divT[row])).real
        H[row, row] += ts2 * (val1 + val2)
        for j in range(row + 1, self.N - 1):
            self.stepper.step(psiH, u[j - 1], u[j], True)
            self.n_steps += 1
            val1 = (of * overlap(self.xiHlist[j], psiH) * normiH).real
            val2 = -(self.divT[row] * np.conj(self.divT[j])).real
            res = ts2 * (val1 + val2)
            H[row, j] += res
            H[j, row] += res

    def calcHessian(self, u, new_control=True, rows=None):
        if new_control:
            self.calculatedXi = False
            self.calcPsiXiDivT(u)
        if not self.calculatedXi:
            self.calcXi(u)
            self.calcDivT(u)
        H = self.calcRegularizationHessian(u)
        of = overlap(self.psi_t[-1], self.psi_target)
        args = self.stepper.args
        for i in range(self.N):
            self.xiHlist[i] = apply_K(self.xi_t[i], args)
        for r in (range(1, self.N - 1) if rows is None else rows):
            self.calcHessianRow(r, u, of, H)
        return H

    def calcFidelityForAllT(self, u, new_control=True):       # :471-491
        if new_control:
            self.calculatedXi = False
            self.calcPsi(u)
        out = []
        for i in range(self.N):
            ov = overlap(self.psi_target, self.psi_t[i])
            out.append(ov.real ** 2 + ov.imag ** 2)
        return out

    # -- public API :495-589
    def propagatePsi(self, control):
        self.calcPsi(control if self.GRAPE else self.basis.convertControl(control))

    def getCost(self, control, new_control=True):
        if self.GRAPE:
            return self.calcCost(control, new_control)
        return self.calcCost(self.basis.convertControl(control, new_control), new_control)

    def getAnalyticGradient(self, control, new_control=True):
        if self.GRAPE:
            return self.calcAnalyticGradient(control, new_control)
        return self.basis.convertGradient(
            self.calcAnalyticGradient(self.basis.convertControl(control, new_control), new_control))

    def getHessian(self, control, new_control=True):
        if self.GRAPE:
            return self.calcHessian(control, new_control)
        return self.basis.convertHessian(
            self.calcHessian(self.basis.convertControl(control, new_control), new_control))

    def getFidelityForAllT(self, control, new_control=True):
        if self.GRAPE:
            return self.calcFidelityForAllT(control, new_control)
        return self.calcFidelityForAllT(self.basis.convertControl(control, new_control), new_control)

    def getControlJacobian(self):
        if self.GRAPE:
            return [[1.0 if i == j else 0.0 for j in range(self.N)] for i in range(self.N)]
        return self.basis.getControlJacobian()
