class AbstractImplicitSolver: pass
class DirectAdjoint: pass
class ODETerm: pass
class SaveAt: pass
def diffeqsolve(*a, **k): raise NotImplementedError
