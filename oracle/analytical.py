"""TEST INFRASTRUCTURE (oracle): the reference's closed-form solutions of the two-phase cylinder
problems, restated.

* `CylindricalSolver`       StefanDG/cylinder-analytical-solution.jl:3-21   (continuous T and flux at r = R)
* `RobinCylindricalSolver`  StefanDG/robin-interface-cylindrical-solution.jl:3-40 (Robin interface:
  T continuous, flux jump k1 T1' - k2 T2' = lambda (Tm - T) at r = R), 3 x 3 system of :23-40.
Core (r < R): T = -q1/(4 k1) r^2 + B1.   Rim: T = -q2/(4 k2) r^2 + A2 log r + B2.
The identities of test_robin_interface_analytical_solution.jl:28-37 and test_analytical_solution.jl are
checked in tests/test_oracle_golden.py."""
import numpy as np


class CylindricalSolver:
    """cylinder-analytical-solution.jl:3-21"""

    def __init__(self, q1, k1, q2, k2, R, b, Tw):
        self.q1, self.k1, self.q2, self.k2, self.R, self.b, self.Tw = q1, k1, q2, k2, R, b, Tw
        self.A2 = (q2 - q1) / (2 * k2) * R ** 2
        self.B2 = Tw + q2 / (4 * k2) * b ** 2 - self.A2 * np.log(b)
        self.B1 = R ** 2 * (q1 / k1 - q2 / k2) / 4 + self.B2 + self.A2 * np.log(R)

    def core_solution(self, r):
        return -self.q1 / (4 * self.k1) * r ** 2 + self.B1

    def rim_solution(self, r):
        return -self.q2 / (4 * self.k2) * r ** 2 + self.A2 * np.log(r) + self.B2

    def __call__(self, r):
        """analytical_solution(r, solver) (:33-39), vectorised"""
        r = np.asarray(r, dtype=float)
        return np.where(r < self.R, self.core_solution(r), self.rim_solution(np.maximum(r, 1e-300)))

    def radial_derivative(self, r):
        """analytical_radial_derivative (:61-69)"""
        r = np.asarray(r, dtype=float)
        core = -self.q1 / (2 * self.k1) * r
        rs = np.maximum(r, 1e-300)
        rim = -self.q1 / (2 * self.k2) * self.R ** 2 / rs - self.q2 / (2 * self.k2) * rs * (1 - self.R ** 2 / rs ** 2)
        return np.where(r < self.R, core, rim)

    def at(self, x, y, center=(0.0, 0.0)):
        return self(np.hypot(np.asarray(x) - center[0], np.asarray(y) - center[1]))


class RobinCylindricalSolver:
    """robin-interface-cylindrical-solution.jl:3-40"""

    def __init__(self, q1, k1, q2, k2, lam, R, b, Tw, Tm):
        self.q1, self.k1, self.q2, self.k2, self.lam, self.R, self.b, self.Tw, self.Tm = q1, k1, q2, k2, lam, R, b, Tw, Tm
        m = self.coefficient_matrix(k2, lam, R, b)
        r = self.coefficient_rhs(q1, k1, q2, k2, lam, R, b, Tw, Tm)
        self.B1, self.A2, self.B2 = np.linalg.solve(m, r)

    @staticmethod
    def coefficient_matrix(k2, lam, R, b):
        """:23-30"""
        return np.array([[1.0, -np.log(R), -1.0], [-lam, k2 / R, 0.0], [0.0, np.log(b), 1.0]])

    @staticmethod
    def coefficient_rhs(q1, k1, q2, k2, lam, R, b, Tw, Tm):
        """:32-39"""
        return np.array([R ** 2 * (q1 / (4 * k1) - q2 / (4 * k2)),
                         0.5 * (q2 - q1) * R - lam * (q1 / (4 * k1) * R ** 2 + Tm),
                         q2 / (4 * k2) * b ** 2 + Tw])

    def core_solution(self, r):
        return -self.q1 / (4 * self.k1) * r ** 2 + self.B1

    def rim_solution(self, r):
        return -self.q2 / (4 * self.k2) * r ** 2 + self.A2 * np.log(r) + self.B2

    def __call__(self, r):
        r = np.asarray(r, dtype=float)
        return np.where(r < self.R, self.core_solution(r), self.rim_solution(np.maximum(r, 1e-300)))

    def core_radial_derivative(self, r):
        return -self.q1 / (2 * self.k1) * r

    def rim_radial_derivative(self, r):
        return -self.q2 / (2 * self.k2) * r + self.A2 / r

    def at(self, x, y, center=(0.0, 0.0)):
        return self(np.hypot(np.asarray(x) - center[0], np.asarray(y) - center[1]))
