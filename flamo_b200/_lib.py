"""ctypes binding of libfsweep.so (include/fsweep.h).  No torch types cross this boundary:
only raw device pointers, sizes and the CUDA stream handle.

The library is built in-tree by `make -C flamo_b200/csrc` (see __graft_entry__.build()).  If it
is missing the import of any compute entry point fails loudly — there is no CPU or PyTorch
fallback for the sweep.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfsweep.so")

OK, E_BADARG, E_UNSUPPORTED, E_WORKSPACE, E_CUDA = 0, -1, -2, -3, -4
C64, C128 = 0, 1
DT_GRAD32 = 256  # FSWEEP_DT_GRAD32: OR-ed into C128 for float32 models swept in float64 arithmetic
EPI_NONE, EPI_ABS = 0, 1
CRIT_MSE, CRIT_MSE_CHSUM = 1, 2
OP_GAIN, OP_PGAIN, OP_SOS, OP_PSOS, OP_DELAY, OP_PDELAY, OP_TABLE, OP_PTABLE, OP_RECURSION = range(1, 10)
F_ISINT, F_GRAD = 1, 2

EXPORTS = [
    "fsweep_version", "fsweep_last_error", "fsweep_plan_create", "fsweep_plan_destroy",
    "fsweep_plan_num_coeffs", "fsweep_plan_coeff_numel", "fsweep_plan_kernel_family", "fsweep_workspace_bytes",
    "fsweep_forward", "fsweep_backward", "fsweep_forward_loss", "fsweep_backward_loss", "fsweep_last_launch_count",
    "fsweep_expm_max_n", "fsweep_expm_forward", "fsweep_expm_backward", "fsweep_expm_forward_sp", "fsweep_expm_backward_sp", "fsweep_expm_backward_sp_total",
    "fsweep_sparsity_forward", "fsweep_sparsity_backward", "fsweep_weighted_total", "fsweep_weighted_total_notify",
    "fsweep_allreduce_p2p", "fsweep_allreduce_p2p_max_n", "fsweep_allreduce_push", "fsweep_allreduce_push_notify", "fsweep_adam_step", "fsweep_adam_step_total", "fsweep_fma_probe", "fsweep_fma_probe_flops", "fsweep_biquad_design", "fsweep_svf_design",
    "fsweep_upload", "fsweep_rfft_supported", "fsweep_rfft_workspace_bytes", "fsweep_rfft_table_entries", "fsweep_rfft_table", "fsweep_rfft",
]


class Op(C.Structure):
    """fsweep_op_t"""
    _fields_ = [
        ("kind", C.c_int32), ("n_out", C.c_int32), ("n_in", C.c_int32), ("n_sections", C.c_int32),
        ("flags", C.c_uint32), ("n_ff", C.c_int32), ("n_fb", C.c_int32), ("per_item", C.c_int32),
    ]


class Seg(C.Structure):
    """fsweep_seg_t"""
    _fields_ = [("ptr", C.c_void_p), ("numel", C.c_int64)]


class Criterion(C.Structure):
    """fsweep_criterion_t"""
    _fields_ = [
        ("kind", C.c_int32), ("reserved", C.c_int32), ("target", C.c_void_p), ("target_batch_stride", C.c_int64),
        ("scale", C.c_double), ("loss", C.c_void_p),
    ]


class AdamTensor(C.Structure):
    """fsweep_adam_tensor_t"""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("step", C.c_void_p), ("numel", C.c_int64)]


ADAM_MAX_TENSORS = 32
MAX_CRITERIA = 8


class TotalJob(C.Structure):
    """fsweep_total_job_t"""
    _fields_ = [("parts", C.c_void_p * MAX_CRITERIA), ("alphas", C.c_double * MAX_CRITERIA),
                ("scales", C.c_double * MAX_CRITERIA), ("n", C.c_int32), ("reserved", C.c_int32), ("vals", C.c_void_p),
                ("host_vals", C.c_void_p), ("host_seq", C.c_void_p), ("seq_counter", C.c_void_p)]


class SweepError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfsweep error {code}: {msg}")
        self.code = code


class Unsupported(SweepError):
    pass


_lib = None


def lib():
    """Load libfsweep.so once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA sweep library is not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C flamo_b200/csrc -j8`. "
            "flamo_b200 has no CPU fallback for the frequency sweep."
        )
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.fsweep_version.restype = i32
    L.fsweep_last_error.restype = C.c_char_p
    L.fsweep_last_launch_count.restype = i32
    L.fsweep_plan_create.restype = i32
    L.fsweep_plan_create.argtypes = [C.POINTER(Op), i32, i64, C.c_double, i32, C.POINTER(vp)]
    L.fsweep_plan_destroy.restype = i32
    L.fsweep_plan_destroy.argtypes = [vp]
    L.fsweep_plan_num_coeffs.restype = i32
    L.fsweep_plan_num_coeffs.argtypes = [vp]
    L.fsweep_plan_coeff_numel.restype = i64
    L.fsweep_plan_coeff_numel.argtypes = [vp, i32, i64]
    L.fsweep_plan_kernel_family.restype = C.c_char_p
    L.fsweep_plan_kernel_family.argtypes = [vp, i64, i32]
    L.fsweep_workspace_bytes.restype = C.c_size_t
    L.fsweep_workspace_bytes.argtypes = [vp, i64, i64, i64]
    L.fsweep_forward.restype = i32
    L.fsweep_forward.argtypes = [vp, C.POINTER(vp), vp, i64, vp, i64, i64, i64, i64, i64, i32, vp]
    L.fsweep_backward.restype = i32
    L.fsweep_backward.argtypes = [vp, C.POINTER(vp), vp, i64, vp, i64, C.POINTER(vp), vp, i64, i64, i64, i64, i64,
                                  i32, vp, C.c_size_t, vp]
    L.fsweep_forward_loss.restype = i32
    L.fsweep_forward_loss.argtypes = [vp, C.POINTER(vp), vp, i64, C.POINTER(Criterion), i64, i64, i64, vp, C.c_size_t, vp]
    L.fsweep_backward_loss.restype = i32
    L.fsweep_backward_loss.argtypes = [vp, C.POINTER(vp), vp, i64, C.POINTER(Criterion), C.POINTER(vp), vp, i64, i64,
                                       i64, i64, vp, C.c_size_t, vp]
    L.fsweep_expm_max_n.restype = i32
    L.fsweep_expm_forward.restype = i32
    L.fsweep_expm_forward.argtypes = [vp, vp, i32, i32, i32, vp]
    L.fsweep_expm_backward.restype = i32
    L.fsweep_expm_backward.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.fsweep_expm_forward_sp.restype = i32
    L.fsweep_expm_forward_sp.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.fsweep_expm_backward_sp.restype = i32
    L.fsweep_expm_backward_sp.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp]
    L.fsweep_expm_backward_sp_total.restype = i32
    L.fsweep_expm_backward_sp_total.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, C.POINTER(TotalJob), vp]
    L.fsweep_sparsity_forward.restype = i32
    L.fsweep_sparsity_forward.argtypes = [vp, i32, i32, i32, vp, vp]
    L.fsweep_sparsity_backward.restype = i32
    L.fsweep_sparsity_backward.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.fsweep_allreduce_p2p_max_n.restype = i32
    L.fsweep_allreduce_p2p.restype = i32
    L.fsweep_allreduce_p2p.argtypes = [vp, vp, i32, i32, i32, C.c_double, vp, vp]
    L.fsweep_allreduce_push.restype = i32
    L.fsweep_allreduce_push.argtypes = [C.POINTER(Seg), i32, vp, vp, i32, i32, i32, C.c_double, vp, vp]
    L.fsweep_allreduce_push_notify.restype = i32
    L.fsweep_allreduce_push_notify.argtypes = [C.POINTER(Seg), i32, vp, vp, i32, i32, i32, C.c_double, vp, vp, vp, vp, vp]
    L.fsweep_adam_step.restype = i32
    L.fsweep_adam_step.argtypes = [C.POINTER(AdamTensor), i32, i32, vp, C.c_double, C.c_double, C.c_double, vp]
    L.fsweep_adam_step_total.restype = i32
    L.fsweep_adam_step_total.argtypes = [C.POINTER(AdamTensor), i32, i32, vp, C.c_double, C.c_double, C.c_double,
                                         C.POINTER(TotalJob), vp]
    L.fsweep_biquad_design.restype = i32
    L.fsweep_biquad_design.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    L.fsweep_svf_design.restype = i32
    L.fsweep_svf_design.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    L.fsweep_fma_probe.restype = i32
    L.fsweep_fma_probe.argtypes = [vp, i32, i32, vp]
    L.fsweep_fma_probe_flops.restype = C.c_double
    L.fsweep_fma_probe_flops.argtypes = [i32, i32]
    L.fsweep_upload.restype = i32
    L.fsweep_upload.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), i32, vp]
    L.fsweep_rfft_supported.restype = i32
    L.fsweep_rfft_supported.argtypes = [i64]
    L.fsweep_rfft_workspace_bytes.restype = C.c_size_t
    L.fsweep_rfft_workspace_bytes.argtypes = [i64, i64]
    L.fsweep_rfft_table_entries.restype = i64
    L.fsweep_rfft_table_entries.argtypes = [i64]
    L.fsweep_rfft_table.restype = i32
    L.fsweep_rfft_table.argtypes = [vp, i64, vp]
    L.fsweep_rfft.restype = i32
    L.fsweep_rfft.argtypes = [vp, i64, i64, i64, i64, i64, C.c_double, vp, vp, vp, C.c_size_t, vp, vp]
    L.fsweep_weighted_total.restype = i32
    L.fsweep_weighted_total.argtypes = [C.POINTER(vp), C.POINTER(C.c_double), C.POINTER(C.c_double), i32, i32, vp, vp]
    L.fsweep_weighted_total_notify.restype = i32
    L.fsweep_weighted_total_notify.argtypes = [C.POINTER(vp), C.POINTER(C.c_double), C.POINTER(C.c_double), i32, i32, vp, vp,
                                               vp, vp, vp]
    _lib = L
    return L


def check(code):
    if code == OK:
        return
    msg = lib().fsweep_last_error().decode()
    raise (Unsupported if code == E_UNSUPPORTED else SweepError)(code, msg)


class Plan:
    """Owns one fsweep_plan_t."""

    def __init__(self, ops, nfft, alias_decay_db, dtype):
        L = lib()
        arr = (Op * len(ops))(*ops)
        h = C.c_void_p()
        check(L.fsweep_plan_create(arr, len(ops), int(nfft), float(alias_decay_db), int(dtype), C.byref(h)))
        self.handle = h
        self.dtype = dtype
        self.n_coeffs = L.fsweep_plan_num_coeffs(h)
        self.launches = 0

    def coeff_numel(self, slot, M):
        return lib().fsweep_plan_coeff_numel(self.handle, slot, M)

    def kernel_family(self, n_bins, backward):
        return lib().fsweep_plan_kernel_family(self.handle, int(n_bins), int(bool(backward))).decode()

    def workspace_bytes(self, batch, cols, n_bins):
        return lib().fsweep_workspace_bytes(self.handle, batch, cols, n_bins)

    def forward(self, coef_ptrs, x_ptr, xbs, y_ptr, ybs, batch, cols, bin_begin, n_bins, epilogue, stream):
        L = lib()
        cp = (C.c_void_p * len(coef_ptrs))(*coef_ptrs)
        check(L.fsweep_forward(self.handle, cp, x_ptr, xbs, y_ptr, ybs, batch, cols, bin_begin, n_bins, epilogue,
                               stream))
        n = L.fsweep_last_launch_count()
        self.launches += n
        return n

    def backward(self, coef_ptrs, x_ptr, xbs, gy_ptr, gybs, grad_ptrs, gx_ptr, gxbs, batch, cols, bin_begin, n_bins,
                 epilogue, ws_ptr, ws_bytes, stream):
        L = lib()
        cp = (C.c_void_p * len(coef_ptrs))(*coef_ptrs)
        gp = (C.c_void_p * len(grad_ptrs))(*grad_ptrs)
        check(L.fsweep_backward(self.handle, cp, x_ptr, xbs, gy_ptr, gybs, gp, gx_ptr, gxbs, batch, cols, bin_begin,
                                n_bins, epilogue, ws_ptr, ws_bytes, stream))
        n = L.fsweep_last_launch_count()
        self.launches += n
        return n

    def forward_loss(self, coef_ptrs, x_ptr, xbs, crit, batch, bin_begin, n_bins, ws_ptr, ws_bytes, stream):
        L = lib()
        cp = (C.c_void_p * len(coef_ptrs))(*coef_ptrs)
        check(L.fsweep_forward_loss(self.handle, cp, x_ptr, xbs, C.byref(crit), batch, bin_begin, n_bins, ws_ptr,
                                    ws_bytes, stream))
        n = L.fsweep_last_launch_count()
        self.launches += n
        return n

    def backward_loss(self, coef_ptrs, x_ptr, xbs, crit, grad_ptrs, gx_ptr, gxbs, batch, bin_begin, n_bins, ws_ptr,
                      ws_bytes, stream):
        L = lib()
        cp = (C.c_void_p * len(coef_ptrs))(*coef_ptrs)
        gp = (C.c_void_p * len(grad_ptrs))(*grad_ptrs)
        check(L.fsweep_backward_loss(self.handle, cp, x_ptr, xbs, C.byref(crit), gp, gx_ptr, gxbs, batch, bin_begin,
                                     n_bins, ws_ptr, ws_bytes, stream))
        n = L.fsweep_last_launch_count()
        self.launches += n
        return n

    def __del__(self):
        try:
            if self.handle:
                lib().fsweep_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
