"""ctypes binding of the C-ABI CUDA library (include/fullbatch_b200.h).

There is no CPU or library fallback: if ``libfullbatch_b200.so`` is missing or a call fails, a ``RuntimeError`` is
raised.  Tensors are passed as raw device pointers (``tensor.data_ptr()``) and the current CUDA stream handle.
"""
import ctypes as C
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfullbatch_b200.so")

FB_TMAP_BYTES = 128
FB_MAX_TAPS = 9
FB_MAX_A_MAPS = 8
FB_MAX_B_MAPS = 2
FB_MAX_WGRAD_TAPS = 9
FB_MAX_GROUPS = 16

vp = C.c_void_p
i32 = C.c_int32
i64 = C.c_int64
f32 = C.c_float


class Tap(C.Structure):
    _fields_ = [("phase", C.c_int8), ("dh", C.c_int8), ("dw", C.c_int8), ("pad", C.c_int8), ("b_k0", i32)]


class TapGroup(C.Structure):
    _fields_ = [("tap0", i32), ("n_taps", i32), ("out_off", i64)]


class ConvGemmArgs(C.Structure):
    _fields_ = [("host_a_maps", vp), ("host_b_maps", vp), ("n_phases", i32), ("a_planes", i32), ("b_planes", i32),
                ("n_taps", i32), ("cblocks", i32), ("taps", Tap * FB_MAX_TAPS), ("tile_w", i32), ("tile_h", i32),
                ("tile_n", i32), ("grid_h", i32), ("grid_n", i32), ("n_total", i32), ("n_tile", i32), ("out", vp),
                ("out_sn", i64), ("out_sh", i64), ("out_sw", i64), ("accumulate", i32), ("n_tapgroups", i32),
                ("tapgroups", TapGroup * 4), ("mg_imgs", i32), ("ng", i32), ("b_group_rows", i32), ("reverse", i32),
                ("stats_ws", vp), ("tickets", vp), ("bn_mean", vp), ("bn_rstd", vp), ("bn_batch", vp), ("bn_eps", f32),
                ("cta_pair", i32), ("halo", i32), ("policy_groups", i32), ("sched_k_iters", i32)]


class WgradTap(C.Structure):
    _fields_ = [("phase", C.c_int8), ("dh", C.c_int8), ("dw", C.c_int8), ("k_index", C.c_int8)]


class WgradArgs(C.Structure):
    _fields_ = [("host_dy_map", vp), ("host_x_maps", vp), ("n_x_maps", i32), ("planes", i32), ("n_taps", i32),
                ("cblocks", i32), ("taps", WgradTap * FB_MAX_WGRAD_TAPS), ("slots_per_cta", i32), ("cout", i32),
                ("cin", i32), ("tile_w", i32), ("tile_h", i32), ("tile_n", i32), ("grid_h", i32), ("grid_n", i32),
                ("splits", i32), ("out", vp), ("out_gstride", i64), ("out_sstride", i64), ("halo", i32),
                ("mg_imgs", i32), ("ng", i32)]


class ReduceEntry(C.Structure):
    _fields_ = [("src", vp), ("src_gstride", i64), ("src_sstride", i64), ("dst_off", i64), ("dst_ld", i64),
                ("splits", i32), ("rows", i32), ("cols", i32), ("src_ld", i32), ("block_start", i32), ("n_blocks", i32),
                ("vec", i32), ("pad", i32)]


class WprepEntry(C.Structure):
    _fields_ = [("w_offset", i64), ("cout", i32), ("cin", i32), ("taps", i32), ("block_start", i32), ("n_blocks", i32),
                ("pad", i32), ("wf_hi", vp), ("wf_lo", vp), ("wd_hi", vp), ("wd_lo", vp), ("ld_f", i64), ("ld_d", i64),
                ("wf_gstride", i64), ("wd_gstride", i64)]


class BnApplyArgs(C.Structure):
    _fields_ = [("y", vp), ("mean", vp), ("rstd", vp), ("gamma", vp), ("beta", vp), ("y2", vp), ("mean2", vp),
                ("rstd2", vp), ("gamma2", vp), ("beta2", vp), ("res_hi", vp), ("res_lo", vp), ("relu", i32), ("P", i64),
                ("C", i32), ("out_hi", vp), ("out_lo", vp), ("ng", i32), ("param_gstride", i64), ("reverse", i32),
                ("mask_out", vp)]


class BnBwdArgs(C.Structure):
    _fields_ = [("dA", vp), ("dA2", vp), ("mask_hi", vp), ("y", vp), ("mean", vp), ("rstd", vp), ("gamma", vp),
                ("P", i64), ("C", i32), ("ws", vp), ("dgamma", vp), ("dbeta", vp), ("dy_bf16", vp), ("dz_out", vp),
                ("ng", i32), ("param_gstride", i64), ("grad_gstride", i64), ("policy_groups", i32), ("reverse", i32),
                ("mask_bits", vp)]


class BnEmaEntry(C.Structure):
    _fields_ = [("running_mean", vp), ("running_var", vp), ("batch", vp), ("pass_stride", i64), ("C", i32),
                ("c_start", i32)]


_SIGNATURES = {
    "fb_version": ([], i32),
    "fb_last_error": ([C.c_char_p, C.c_size_t], i32),
    "fb_tmap_encode_act4d": ([vp, vp, i32, i32, i32, i32, i64, i64, i64, i32, i32, i32, i32], i32),
    "fb_tmap_encode_mat2d": ([vp, vp, i32, i32, i64, i32, i32], i32),
    "fb_conv_gemm": ([C.POINTER(ConvGemmArgs), vp], i32),
    "fb_conv_stats_rows": ([i32, i32, i32, i32], i32),
    "fb_conv_pair_ok": ([i32, i32, i32, i32], i32),
    "fb_conv_wgrad": ([C.POINTER(WgradArgs), vp], i32),
    "fb_reduce_multi": ([vp, i32, i32, vp, i64, i32, vp], i32),
    "fb_weight_prep": ([vp, i32, i32, i32, vp, vp, i64, vp, vp, i64, vp], i32),
    "fb_weight_prep_multi": ([vp, vp, i32, i32, vp, i64, vp, f32, f32, f32, vp, i32, i32, vp], i32),
    "fb_stem_im2col": ([vp, vp, vp, vp, i64, i32, i32, vp, vp, vp, vp], i32),
    "fb_stem_im2col_u8aug": ([vp, vp, vp, vp, i64, i32, i32, vp, C.POINTER(f32), C.POINTER(f32), vp, vp, vp, vp], i32),
    "fb_bn_stats": ([vp, i64, i32, vp, vp, vp, vp, vp, f32, f32, vp], i32),
    "fb_bn_apply": ([C.POINTER(BnApplyArgs), vp], i32),
    "fb_bn_bwd": ([C.POINTER(BnBwdArgs), vp], i32),
    "fb_bn_bwd_chunks": ([i64, i32, i32], i32),
    "fb_bn_ema_multi": ([vp, i32, i32, i32, i32, f32, vp], i32),
    "fb_avgpool2_fwd": ([vp, vp, i32, i32, i32, i32, vp, vp, vp], i32),
    "fb_avgpool2_bwd": ([vp, i32, i32, i32, i32, vp, i32, vp], i32),
    "fb_head_fwd_bwd": ([vp, vp, i32, i32, i32, vp, vp, vp, i32, f32, vp, vp, i32, i32, vp, vp, vp, i32, i64, i64, vp],
                        i32),
    "fb_flat_sqnorm": ([vp, i64, vp, f32, f32, i64, i32, vp, vp, i32, vp, vp, i32, f32, f32, i32, vp], i32),
    "fb_perturb_ranges": ([vp, vp, i64, vp, vp, i32, i64, f32, f32, f32, vp, i32, vp, i64, i32, vp], i32),
    "fb_fd_combine": ([vp, vp, vp, i64, vp, i64, i32, vp, i32, i32, vp, i32, vp], i32),
    "fb_mean_accumulate": ([vp, i64, vp, i64, i32, vp, vp, i32, f32, i32, vp], i32),
    "fb_group_finish": ([vp, i32, i32, vp, vp, i32, i32, i32, i32, vp], i32),
    "fb_flat_scale": ([vp, i64, f32, vp], i32),
    "fb_flat_relayout": ([vp, vp, i64, vp, i32, i32, i32, vp], i32),
    "fb_sgd_step": ([vp, vp, vp, i64, vp, i32, f32, f32, f32, f32, f32, i32, i32, i32, vp, i32, vp], i32),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None


def load(build_if_missing=True):
    """Load (building first if the sources are newer and nvcc is available) and return the ctypes library."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and _build.needs_build() and os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
        _build.build_library()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the library does not export what the header declares
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def last_error():
    buf = C.create_string_buffer(512)
    load().fb_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(code, what):
    if code != 0:
        raise RuntimeError(f"{what} failed with code {code}: {last_error()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def stream_handle():
    import torch

    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Call an exported function on the current stream; raises RuntimeError on failure."""
    fn = getattr(load(), name)
    check(fn(*args, stream_handle()), name)
