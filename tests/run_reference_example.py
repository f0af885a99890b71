"""Runs ONE unmodified example script of the reference (`/root/reference/examples/*.py`) in this process, either on the
reference itself or on flamo_b200 installed under the name `flamo` (install_as_flamo) with the C ABI emulated on the
CPU (tests/cpu_emulator.py: the build container has no GPU).  Used by tests/test_reference_examples_cpu.py, one
subprocess per run so that the two `flamo` packages never share an interpreter.

    python tests/run_reference_example.py {reference|b200} <example.py> <train_dir> [script args...]
"""
import os
import runpy
import sys
import types
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    engine, script, train_dir = sys.argv[1:4]
    extra = sys.argv[4:]
    # third-party packages the example scripts import but this image does not have: plotting and audio file IO
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = mock.MagicMock(name=name)
                m.subplots = lambda *a, **k: (mock.MagicMock(), (mock.MagicMock(), mock.MagicMock()))
                sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if engine == "reference":
        for name in ("soundfile", "nnAudio", "nnAudio.features", "pyfar"):
            if name not in sys.modules:
                try:
                    __import__(name)
                except Exception:
                    sys.modules[name] = types.ModuleType(name)
        sys.modules["nnAudio"].features = sys.modules["nnAudio.features"]
        if not hasattr(sys.modules["soundfile"], "write"):
            sys.modules["soundfile"].write = lambda path, data, fs, subtype=None: open(path, "wb").close()
        sys.path.insert(0, "/root/reference")
    else:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, HERE)
        import flamo_b200

        flamo_b200.install_as_flamo()
        import cpu_emulator

        cpu_emulator.install()
    sys.argv = [script, "--train_dir", train_dir] + extra
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
