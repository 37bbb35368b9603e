"""ctypes binding of include/dregb200.h.  There is NO fallback: a missing library or a failing
call raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DRB_LIB_PATH") or os.path.join(HERE, "libdregb200.so")

c_void_p, c_int, c_ll, c_float, c_size_t = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t


class DrbError(RuntimeError):
    pass


class Conv3dDesc(C.Structure):
    _fields_ = [("g", c_int), ("d", c_int), ("h", c_int), ("w", c_int),
                ("cin", c_int), ("cout", c_int),
                ("kd", c_int), ("kh", c_int), ("kw", c_int),
                ("planes", c_int), ("relu", c_int), ("acc_scale", c_float), ("out_scale", c_float),
                ("x_hi", c_void_p), ("x_lo", c_void_p), ("w_hi", c_void_p), ("w_lo", c_void_p),
                ("bias", c_void_p), ("residual", c_void_p),
                ("out", c_void_p), ("out_hi", c_void_p), ("out_lo", c_void_p),
                ("ld_out", c_ll), ("bn_accum", c_void_p), ("tile_list", c_void_p), ("tile_count", c_void_p),
                ("acc_scale_dev", c_void_p * 2), ("splitk_ws", c_void_p), ("splitk_ws_bytes", c_size_t),
                ("res_d", c_int), ("res_h", c_int), ("res_w", c_int)]


class WgradDesc(C.Structure):
    _fields_ = [("g", c_int), ("d", c_int), ("h", c_int), ("w", c_int), ("cout", c_int), ("cin", c_int),
                ("kd", c_int), ("kh", c_int), ("kw", c_int), ("planes", c_int),
                ("dy_hi", c_void_p), ("dy_lo", c_void_p), ("x_hi", c_void_p), ("x_lo", c_void_p),
                ("scale", c_float), ("scale_dev", c_void_p), ("dw", c_void_p),
                ("c_real", c_int), ("taps_real", c_int), ("tile_list", c_void_p), ("tile_count", c_void_p),
                ("stage", c_void_p), ("stage_elems", c_ll)]


class PairGrad(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("d_src_feats", "d_tgt_feats", "d_src_corr", "d_tgt_corr",
                                        "d_src_overlap", "d_tgt_overlap", "d_pose")]


class Im2colDesc(C.Structure):
    _fields_ = [("x", c_void_p), ("sg", c_ll), ("sc", c_ll), ("sd", c_ll), ("sh", c_ll), ("sw", c_ll),
                ("g", c_int), ("c", c_int), ("d", c_int), ("h", c_int), ("w", c_int),
                ("k", c_int), ("stride", c_int), ("pad", c_int), ("kpad", c_int)]


class NgpParams(C.Structure):
    _fields_ = [("hash_table", c_void_p), ("w1", c_void_p), ("w2", c_void_p),
                ("c1", c_void_p), ("c2", c_void_p), ("c3", c_void_p), ("aabb", c_float * 6)]


class ExtractDesc(C.Structure):
    _fields_ = [("res", c_int), ("roi_aabb", c_float * 6), ("scene_aabb", c_float * 6),
                ("occupied", c_void_p), ("n_occupied", c_int), ("jitter", c_void_p),
                ("occ_binary", c_void_p), ("cam_origins", c_void_p), ("ncams", c_int),
                ("render_step_size", c_float), ("density_thre", c_float), ("cut_off", c_float),
                ("host_dirs", C.POINTER(c_float)), ("ndirs", c_int), ("surface_only_where_dense", c_int),
                ("rgb_only_where_masked", c_int)]


class EngineConfig(C.Structure):
    _fields_ = [("res_x", c_int), ("res_y", c_int), ("res_z", c_int), ("planes", c_int),
                ("num_downsample", c_int), ("pos_emb_scaling", c_float), ("max_mask", c_int),
                ("training_bn", c_int)]


class PairIO(C.Structure):
    _fields_ = [("src_grid", c_void_p), ("tgt_grid", c_void_p),
                ("s_ch", c_ll), ("s_z", c_ll), ("s_x", c_ll), ("s_y", c_ll),
                ("t_ch", c_ll), ("t_z", c_ll), ("t_x", c_ll), ("t_y", c_ll),
                ("src_mask", c_void_p), ("n_src_mask", c_int),
                ("tgt_mask", c_void_p), ("n_tgt_mask", c_int)]


class PairOut(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("src_feats", "tgt_feats", "src_kp", "tgt_kp", "src_corr",
                                        "tgt_corr", "src_overlap", "tgt_overlap", "pose")]


# name -> (restype, argtypes).  Every symbol declared in include/dregb200.h is listed here; the CPU
# test-suite checks that the shared library exports all of them.
SIGNATURES = {
    "drb_abi_version": (c_int, []),
    "drb_last_error": (C.c_char_p, []),
    "drb_igemm_error_flag": (c_int, [C.POINTER(c_int)]),
    "drb_error_flag_clear": (c_int, []),
    "drb_error_flag_peek": (c_int, [C.POINTER(c_int)]),
    "drb_error_flag_detail": (c_int, [C.POINTER(c_int * 16)]),
    "drb_conv3d_igemm": (c_int, [C.POINTER(Conv3dDesc), c_void_p]),
    "drb_conv3d_tile_shape": (c_int, [c_int, c_int, c_int, c_int, C.POINTER(c_int * 4), C.POINTER(c_int * 4)]),
    "drb_split_planes": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "drb_weight_scale": (c_int, [c_void_p, c_ll, C.POINTER(c_float), c_void_p]),
    "drb_pack_conv_weight": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "drb_pack_conv_weight_im2col": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                            c_void_p]),
    "drb_im2col": (c_int, [C.POINTER(Im2colDesc), c_void_p, c_void_p, c_void_p]),
    "drb_im2col_stem": (c_int, [c_void_p, c_ll, c_ll, c_ll, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "drb_bn_stats": (c_int, [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p]),
    "drb_bn_small_supported": (c_int, [c_ll, c_int]),
    "drb_bn_small": (c_int, [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float,
                             c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p]),
    "drb_bn_finalize": (c_int, [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "drb_scale_shift_act": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_bn_apply": (c_int, [c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                             c_float, c_float, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_maxpool3d": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_upsample2_add": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_trilinear_gather": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_ll, c_ll, c_ll,
                                     c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "drb_fpn_need_tiles": (c_int, [C.POINTER(c_void_p), C.POINTER(c_int), c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_downsample_workspace_bytes": (c_size_t, [c_int, c_int]),
    "drb_hierarchical_downsample": (c_int, [c_void_p, c_int, c_int, c_int, c_int, C.c_double, c_int, c_void_p,
                                            c_size_t, c_void_p, C.POINTER(c_int), C.POINTER(c_int), c_void_p]),
    "drb_pos_embed_sine": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "drb_layernorm256": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "drb_mha_core": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float,
                             c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "drb_mha_tc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "drb_mha_tc_pack": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                c_void_p, c_size_t, c_void_p]),
    "drb_mha_tc_forward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_int, c_void_p]),
    "drb_engine_set_tc_attention": (c_int, [c_void_p, c_int]),
    "drb_softmax_weighted_xyz": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "drb_overlap_sigmoid": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_procrustes": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int,
                               c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int,
                               c_int, c_int, c_void_p, c_void_p]),
    "drb_ngp_table_entries": (c_ll, []),
    "drb_ngp_density": (c_int, [C.POINTER(NgpParams), c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "drb_ngp_rgb_mean": (c_int, [C.POINTER(NgpParams), c_void_p, c_int, C.POINTER(c_float), c_int, c_void_p,
                                 c_void_p]),
    "drb_surface_mask": (c_int, [C.POINTER(NgpParams), c_void_p, c_int, C.POINTER(c_float), C.POINTER(c_float),
                                 c_void_p, c_int, c_void_p, c_int, c_float, c_float, c_void_p, c_void_p]),
    "drb_extract_block": (c_int, [C.POINTER(NgpParams), C.POINTER(ExtractDesc), c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_surface_mask_workspace_bytes": (C.c_size_t, [c_int]),
    "drb_surface_mask_ws": (c_int, [C.POINTER(NgpParams), c_void_p, c_int, C.POINTER(c_float), C.POINTER(c_float),
                                    c_void_p, c_int, c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, C.c_size_t,
                                    c_void_p]),
    "drb_extract_workspace_bytes": (C.c_size_t, [c_int]),
    "drb_extract_block_ws": (c_int, [C.POINTER(NgpParams), C.POINTER(ExtractDesc), c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, C.c_size_t, c_void_p]),
    "drb_extract_last_surface_ms": (c_int, [C.POINTER(c_float)]),
    "drb_march_stats": (c_int, [C.POINTER(C.c_ulonglong), c_int]),
    "drb_extract_set_profile": (c_int, [c_int]),
    "drb_extract_read_profile": (c_int, [C.POINTER(c_float), C.POINTER(c_int)]),
    "drb_conv3d_wgrad": (c_int, [C.POINTER(WgradDesc), c_void_p]),
    "drb_grad_split": (c_int, [c_void_p, c_ll, c_int, c_ll, c_ll, c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_add_inplace": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "drb_relu_mask_plane": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "drb_colsum_add": (c_int, [c_void_p, c_ll, c_int, c_ll, c_void_p, c_void_p]),
    "drb_bn_save_stats": (c_int, [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_bn_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_int, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_maxpool3d_backward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "drb_upsample2_add_backward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                           c_void_p]),
    "drb_trilinear_gather_backward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                              c_void_p, c_int, c_void_p, c_void_p]),
    "drb_col2im": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                           c_void_p]),
    "drb_layernorm256_backward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                          c_void_p]),
    "drb_overlap_sigmoid_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p]),
    "drb_softmax_rows": (c_int, [c_void_p, c_ll, c_int, c_int, c_float, c_void_p]),
    "drb_softmax_backward_rows": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_float, c_void_p]),
    "drb_softmax_weighted_xyz_backward": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "drb_sgemm_strided": (c_int, [c_void_p, c_ll, c_ll, c_ll, c_void_p, c_ll, c_ll, c_ll, c_void_p, c_ll, c_ll, c_int,
                                  c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "drb_mha_backward_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "drb_mha_core_backward": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                      c_int, c_float, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                      c_size_t, c_void_p]),
    "drb_procrustes_backward": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int,
                                        c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "drb_segment_mean_backward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "drb_hierarchical_downsample_tape": (c_int, [c_void_p, c_int, c_int, c_int, c_int, C.c_double, c_int, c_void_p,
                                                 c_size_t, c_void_p, C.POINTER(c_int), C.POINTER(c_int), c_void_p, c_ll,
                                                 C.POINTER(c_int), C.POINTER(c_int), c_void_p]),
    "drb_fpn_dilated_tiles": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    "drb_adamw_create": (c_int, [c_int, C.POINTER(c_void_p), C.POINTER(c_ll), C.POINTER(c_void_p)]),
    "drb_adamw_destroy": (None, [c_void_p]),
    "drb_adamw_step": (c_int, [c_void_p, C.POINTER(c_void_p), c_float, c_float, c_float, c_float, c_float, c_float,
                               c_void_p]),
    "drb_adamw_grad_norm": (c_int, [c_void_p, C.POINTER(C.c_double), c_void_p]),
    "drb_adamw_copy_state": (c_int, [c_void_p, c_int, c_void_p, c_void_p, C.POINTER(c_ll), c_void_p]),
    "drb_adamw_set_step": (c_int, [c_void_p, c_ll]),
    "drb_engine_set_grad_mode": (c_int, [c_void_p, c_int]),
    "drb_engine_param_trainable": (c_int, [c_void_p, c_int]),
    "drb_engine_bind_grad": (c_int, [c_void_p, c_int, c_void_p]),
    "drb_engine_backward": (c_int, [c_void_p, C.POINTER(PairIO), C.POINTER(PairOut), C.POINTER(PairGrad), c_void_p]),
    "drb_engine_set_max_tokens": (c_int, [c_void_p, c_int]),
    "drb_engine_create": (c_int, [C.POINTER(EngineConfig), C.POINTER(c_void_p)]),
    "drb_engine_destroy": (None, [c_void_p]),
    "drb_engine_num_params": (c_int, [c_void_p]),
    "drb_engine_param_name": (C.c_char_p, [c_void_p, c_int]),
    "drb_engine_param_numel": (c_ll, [c_void_p, c_int]),
    "drb_engine_bind_param": (c_int, [c_void_p, c_int, c_void_p]),
    "drb_engine_commit_params": (c_int, [c_void_p, c_void_p]),
    "drb_engine_set_training": (c_int, [c_void_p, c_int]),
    "drb_engine_set_sparse_fpn": (c_int, [c_void_p, c_int]),
    "drb_engine_set_update_running": (c_int, [c_void_p, c_int]),
    "drb_engine_encode": (c_int, [c_void_p, C.POINTER(PairIO), C.POINTER(c_int), C.POINTER(c_int), c_void_p]),
    "drb_engine_decode": (c_int, [c_void_p, C.POINTER(PairOut), c_void_p]),
    "drb_engine_tap": (c_int, [c_void_p, C.c_char_p, c_int, c_void_p, c_ll, C.POINTER(c_ll), c_void_p]),
    "drb_engine_launch_count": (c_ll, [c_void_p]),
    "drb_engine_set_profile": (c_int, [c_void_p, c_int]),
    "drb_engine_profile_read": (c_int, [c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(c_ll)]),
}

_lib = None


def load():
    """Loads libdregb200.so (building nothing: see build.py / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DrbError("libdregb200.so is missing at %s - run `python __graft_entry__.py build`; "
                       "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.drb_abi_version() != 4:
        raise DrbError("libdregb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().drb_last_error()
        raise DrbError("%s failed with code %d: %s" % (what or "libdregb200 call", rc,
                                                      msg.decode(errors="replace") if msg else ""))


_FLAG_TEXT = {21: "a mask index is outside the grid", 31: "the surface-field marcher hit its watchdog: result truncated"}


def check_device_flag(what=""):
    """Reads the current device's sticky error flag (synchronises); raises and clears it when set."""
    lib = load()
    v = C.c_int(0)
    check(lib.drb_igemm_error_flag(C.byref(v)), "drb_igemm_error_flag")
    if v.value != 0:
        lib.drb_error_flag_clear()
        raise DrbError("%s: device error flag %d (%s)" % (what or "libdregb200", v.value,
                                                          _FLAG_TEXT.get(v.value, "tensor-core pipeline watchdog")))


def check_stream_flag(what=""):
    """Same check for the work queued on the CURRENT stream only: waits for that stream, then reads the flag without
    a device-wide wait - pairs in flight on other streams (pipeline.PairPipeline) keep running."""
    import torch
    lib = load()
    torch.cuda.current_stream().synchronize()
    v = C.c_int(0)
    check(lib.drb_error_flag_peek(C.byref(v)), "drb_error_flag_peek")
    if v.value != 0:
        lib.drb_error_flag_clear()
        raise DrbError("%s: device error flag %d (%s)" % (what or "libdregb200", v.value,
                                                          _FLAG_TEXT.get(v.value, "tensor-core pipeline watchdog")))


def ptr(t):
    """Device (or host) address of a torch tensor, None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
