"""SolverEMI with the reference's interface (src/knpemidg/solver_emi.py:52-822): the EMI
sub-problem alone.  The potential is solved every step with the conductivity
kappa = F psi sum_k z_k^2 D_k c_k of the INITIAL concentrations (solver_emi.py:246-268); the
concentrations are never advanced (there is no KNP step, :491-509), so Nernst potentials and
the eliminated ion stay at their initial values too.

Same device path as `Solver`: knp_assemble_emi + knp_solve_emi + knp_post_step(PHIM) per step,
the membrane ODE step in front of it in `solve_system_active`.
"""
from __future__ import annotations

from . import _lib
from .solver import Solver


class SolverEMI(Solver):
    def solve_for_time_step(self, k, t):
        """one global step (solver_emi.py:491-509): EMI solve, membrane potential, time"""
        eng = self.engine
        self._t = t
        if self.mms is not None:
            self._update_loads()
        ctx = eng.ctx
        ctx.assemble_emi()
        it, _ = ctx.solve_emi(eng.rtol_emi, eng.atol_emi, eng.max_it)
        eng.stats["emi_niter"].append(it)
        eng.stats["knp_niter"].append(0)
        ctx.post_step(_lib.POST_PHIM)
        eng.t += eng.dt
        t.assign(float(t) + float(self.dt))

    def solve_for_time_step_picard(self, k, t):
        """solver_emi.py:512-556: with frozen concentrations the Picard loop is a single EMI solve"""
        t.assign(float(t) + float(self.dt))
        self._t = t
        eng, ctx = self.engine, self.engine.ctx
        ctx.assemble_emi()
        it, _ = ctx.solve_emi(eng.rtol_emi, eng.atol_emi, eng.max_it)
        eng.stats["emi_niter"].append(it)
        eng.stats["knp_niter"].append(0)
        ctx.post_step(_lib.POST_PHIM)
        eng.t += eng.dt
