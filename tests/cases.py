"""Small parity cases shared by the golden-vector generator, the CPU tests and the GPU tests.

Each case is a sub-sweep of a BASELINE configuration (same generators, same options, fewer
instances / phases) that the CPU checker finishes in about a second."""
import numpy as np


def subsample(wl, n, seed=None):
    """Every k-th instance of the full sweep (keeps first and last), as a new workload."""
    import copy
    idx = np.unique(np.linspace(0, wl.NT - 1, n).round().astype(np.int64))
    out = copy.copy(wl)
    out.buffers = wl.buffers.take(idx)
    return out


def make_cases(W):
    """name -> (workload, phases).  W is the product's ``workloads`` module."""
    c = {}
    c["lorenz_rk4_f64"] = (subsample(W.lorenz(NT=1 << 14, algorithm=W.RK4), 96), 1)
    c["lorenz_rkck45_f64"] = (subsample(W.lorenz(NT=1 << 14, algorithm=W.RKCK45), 192), 1)
    c["lorenz_rkck45_f32"] = (subsample(W.lorenz(NT=1 << 14, algorithm=W.RKCK45, dtype=np.float32), 96), 1)
    c["poincare_rk4_f64"] = (subsample(W.poincare_t3(NT=4096, algorithm=W.RK4), 96), 4)
    c["poincare_rkck45_f64"] = (subsample(W.poincare_t3(NT=4096, algorithm=W.RKCK45), 96), 4)
    c["duffing_rk4_f64"] = (subsample(W.duffing_t1(NT=4096, algorithm=W.RK4), 96), 3)
    c["duffing_rkck45_f64"] = (subsample(W.duffing_t1(NT=4096, algorithm=W.RKCK45), 96), 3)
    c["impact_rkck45_f64"] = (subsample(W.impact_t6(NT=1 << 14), 128), 6)
    c["kellermiksis_rkck45_f64"] = (subsample(W.keller_miksis_t5(NT=4096, ndo=0), 64), 2)
    c["kellermiksis_dense_rkck45_f64"] = (subsample(W.keller_miksis_t5(NT=4096, ndo=200), 64), 2)
    return c


RESULT_FIELDS = ("TimeDomain", "ActualState", "ActualTime", "ControlParameters", "Accessories",
                 "IntegerAccessories", "DenseIndex", "DenseTime", "DenseState")
