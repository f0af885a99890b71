"""CPU: the host-side planner of libfsweep.so (fsweep_plan_create and the plan queries run without a GPU) under
random and adversarial op programs.  The contract (include/fsweep.h): every call returns FSWEEP_OK or a negative
error code with a message — never a crash, never an exception across the ABI — and an accepted plan answers its
queries consistently.  Argument checks of the launch entry points are covered with NULL / out-of-range arguments
(they must be rejected before anything is enqueued, so no GPU is needed)."""
import ctypes as C

from hypothesis import HealthCheck, given, settings, strategies as st

from flamo_b200 import _lib

KINDS = [_lib.OP_GAIN, _lib.OP_PGAIN, _lib.OP_SOS, _lib.OP_PSOS, _lib.OP_DELAY, _lib.OP_PDELAY, _lib.OP_TABLE,
         _lib.OP_PTABLE, _lib.OP_RECURSION]
DIAG = {_lib.OP_PGAIN, _lib.OP_PSOS, _lib.OP_PDELAY, _lib.OP_PTABLE}
CODES = {_lib.OK, _lib.E_BADARG, _lib.E_UNSUPPORTED}

small = st.one_of(st.integers(min_value=-2, max_value=70), st.sampled_from([1 << 15, 1 << 30, (1 << 31) - 1]))
op_raw = st.tuples(st.sampled_from(KINDS + [0, 10, -1, 1 << 30]), small, small, st.one_of(st.integers(-1, 40), st.sampled_from([1 << 20, (1 << 31) - 1])),
                   st.integers(0, 7), st.integers(-1, 6), st.integers(-1, 6), st.just(0))


def create(ops, nfft, alias, dtype):
    L = _lib.lib()
    arr = (_lib.Op * max(1, len(ops)))(*[_lib.Op(*o) for o in ops])
    h = C.c_void_p()
    rc = L.fsweep_plan_create(arr, len(ops), int(nfft), float(alias), int(dtype), C.byref(h))
    return rc, h


def query(h, M=1025):
    L = _lib.lib()
    n = L.fsweep_plan_num_coeffs(h)
    assert 0 <= n <= 64
    for s in range(n):
        assert L.fsweep_plan_coeff_numel(h, s, M) > 0
    assert L.fsweep_plan_coeff_numel(h, n, M) <= 0 and L.fsweep_plan_coeff_numel(h, -1, M) <= 0  # out of range slots
    for bwd in (0, 1):
        for nb in (1, 100, 48001):
            assert L.fsweep_plan_kernel_family(h, nb, bwd).decode().startswith("fsweep")
    w1, w2 = L.fsweep_workspace_bytes(h, 1, 1, 1000), L.fsweep_workspace_bytes(h, 4, 2, 1000)
    assert 0 <= w1 <= w2 < (1 << 40)
    assert L.fsweep_plan_destroy(h) == _lib.OK


@settings(max_examples=400, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(op_raw, min_size=0, max_size=12), st.sampled_from([0, 1, 2, 64, 4096, 96000, -8]),
       st.sampled_from([0.0, 30.0, -30.0, float("nan")]), st.sampled_from([_lib.C64, _lib.C128, 5]))
def test_random_programs_never_crash(ops, nfft, alias, dtype):
    rc, h = create(ops, nfft, alias, dtype)
    assert rc in CODES, rc
    if rc == _lib.OK:
        assert h.value
        query(h)
    else:
        assert not h.value and _lib.lib().fsweep_last_error()  # no plan leaks out of a failed create, a message is set


@st.composite
def valid_program(draw):
    """A chain with consistent channel widths and at most one recursion (what the lowering produces)."""
    width = draw(st.integers(1, 16))
    ops = []

    def leaf(n_in, force_square=False):
        kind = draw(st.sampled_from([k for k in KINDS if k != _lib.OP_RECURSION]))
        K = draw(st.integers(1, 6)) if kind in (_lib.OP_SOS, _lib.OP_PSOS) else 0
        flags = draw(st.integers(0, 3))
        n_out = n_in if (kind in DIAG or force_square) else draw(st.integers(1, 16))
        return (kind, n_out, n_in, K, flags, 0, 0, 0), n_out

    n = width
    for _ in range(draw(st.integers(0, 2))):
        o, n = leaf(n)
        ops.append(o)
    if draw(st.booleans()):
        ff, cur = [], n
        for _ in range(draw(st.integers(1, 3))):
            o, cur = leaf(cur)
            ff.append(o)
        N = cur
        fb = []
        for j in range(draw(st.integers(1, 2))):
            o, cur = leaf(cur)
            fb.append(o)
        if cur != n:  # close the loop: the feedback chain must return to the loop input width
            fb.append((_lib.OP_GAIN, n, cur, 0, draw(st.integers(0, 3)), 0, 0, 0))
        ops.append((_lib.OP_RECURSION, N, n, 0, 0, len(ff), len(fb), 0))
        ops += ff + fb
        n = N
    for _ in range(draw(st.integers(0 if ops else 1, 2))):
        o, n = leaf(n)
        ops.append(o)
    return ops


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(valid_program(), st.sampled_from([64, 4096, 96000]), st.sampled_from([_lib.C64, _lib.C128]))
def test_well_formed_programs_are_planned(ops, nfft, dtype):
    rc, h = create(ops, nfft, 30.0, dtype)
    assert rc == _lib.OK, (_lib.lib().fsweep_last_error().decode(), ops)
    assert _lib.lib().fsweep_plan_num_coeffs(h) == sum(o[0] != _lib.OP_RECURSION for o in ops)
    query(h)


def test_launch_entry_points_reject_bad_arguments_before_launching():
    L = _lib.lib()
    rc, h = create([(_lib.OP_GAIN, 2, 2, 0, 2, 0, 0, 0)], 64, 0.0, _lib.C64)
    assert rc == _lib.OK
    null = C.c_void_p()
    one = (C.c_void_p * 1)(None)
    assert L.fsweep_forward(None, one, null, 0, null, 0, 1, 1, 0, 33, 0, null) == _lib.E_BADARG
    assert L.fsweep_forward(h, one, null, 0, null, 0, 1, 1, 0, 33, 0, null) == _lib.E_BADARG       # NULL x / y
    assert L.fsweep_backward(h, one, null, 0, null, 0, one, null, 0, 1, 1, 0, 33, 0, null, 0, null) == _lib.E_BADARG
    assert L.fsweep_plan_create(None, 1, 64, 0.0, _lib.C64, C.byref(C.c_void_p())) == _lib.E_BADARG
    assert L.fsweep_plan_create((_lib.Op * 1)(), 1, 64, 0.0, _lib.C64, None) == _lib.E_BADARG
    assert L.fsweep_expm_forward(None, None, 4, 1, _lib.C64, None) == _lib.E_BADARG
    assert L.fsweep_expm_forward(1, 1, 0, 1, _lib.C64, None) == _lib.E_BADARG
    assert L.fsweep_adam_step(None, 1, _lib.C64, None, 0.9, 0.999, 1e-8, None) == _lib.E_BADARG
    t = (_lib.AdamTensor * 1)(_lib.AdamTensor(1, 1, 1, 1, 1, 4))
    assert L.fsweep_adam_step(t, 0, _lib.C64, 1, 0.9, 0.999, 1e-8, None) == _lib.E_BADARG          # n out of range
    assert L.fsweep_adam_step(t, _lib.ADAM_MAX_TENSORS + 1, _lib.C64, 1, 0.9, 0.999, 1e-8, None) == _lib.E_BADARG
    assert L.fsweep_adam_step(t, 1, 7, 1, 0.9, 0.999, 1e-8, None) == _lib.E_BADARG                 # unknown dtype
    assert L.fsweep_plan_destroy(h) == _lib.OK
    assert L.fsweep_plan_destroy(None) in (_lib.OK, _lib.E_BADARG)
