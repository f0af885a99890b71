"""flamo_b200 — B200-native frequency-sampling engine behind the flamo class API.

Scope: the per-bin complex forward/backward sweep through
processor.dsp.{Gain, parallelGain, Matrix, Filter, Biquad, SVF, GEQ, Delay} (+ parallel variants)
composed by processor.system.{Series, Recursion} inside system.Shell, driven by
optimize.trainer.Trainer — evaluated by hand-written sm_100a kernels in libfsweep.so
(include/fsweep.h).  There is no CPU implementation of the sweep in this package.
"""
__version__ = "0.1.0"


def install_as_flamo():
    """Register this package under the name `flamo` so that unmodified reference scripts
    (`from flamo.processor import dsp, system`, `from flamo.optimize.trainer import Trainer`, ...)
    import the B200 engine instead."""
    import importlib
    import sys

    names = ["", ".utils", ".functional", ".processor", ".processor.dsp", ".processor.system", ".optimize",
             ".optimize.trainer", ".optimize.dataset", ".optimize.loss", ".auxiliary", ".auxiliary.eq"]
    for n in names:
        sys.modules["flamo" + n] = importlib.import_module(__name__ + n)
