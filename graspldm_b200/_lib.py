"""ctypes binding of libgraspldm_b200.so (the C ABI of include/graspldm_b200.h).

The library is built in-tree by graspldm_b200/build.py (nvcc, sm_100a).  Loading fails loudly:
there is no CPU or PyTorch fallback for any entry point.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_longlong, c_ulonglong, c_void_p, c_char_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgraspldm_b200.so")

P = c_void_p  # device / host pointers travel as integers


class GldmSamplerArgs(Structure):
    _fields_ = [("x_init", c_void_p), ("z_obj", c_void_p), ("n", c_int), ("grasps_per_obj", c_int), ("sched_kind", c_int),
                ("n_steps", c_int), ("coef", c_void_p), ("timesteps", c_void_p), ("times", c_void_p), ("te", c_void_p),
                ("clip_sample", c_int), ("noise", c_void_p), ("seed", c_ulonglong), ("cls_emb", c_void_p),
                ("x_out", c_void_p), ("x_all", c_void_p)]


class GldmResNetCfg(Structure):
    _fields_ = [("L", c_int), ("n_stages", c_int), ("ch", c_int * 6), ("emb_dim", c_int), ("cond_ch", c_int),
                ("cond_dim", c_int), ("groups", c_int), ("time_cond", c_int), ("fourier_half", c_int),
                ("heads", c_int), ("dim_head", c_int)]


# name -> argtypes (all return int unless listed in _RESTYPES)
_SIGS = {
    "gldm_avg_voxelize_forward": [P, P, c_int, c_int, c_int, c_int, P, P, P, P],
    "gldm_avg_voxelize_backward": [P, P, P, c_int, c_int, c_int, c_int, P, P],
    "gldm_trilinear_devoxelize_forward": [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P],
    "gldm_trilinear_devoxelize_backward": [P, P, P, c_int, c_int, c_int, c_int, P, P],
    "gldm_furthest_point_sampling": [P, c_int, c_int, c_int, P, P],
    "gldm_gather_features_forward": [P, P, c_int, c_int, c_int, c_int, P, P],
    "gldm_gather_features_backward": [P, P, c_int, c_int, c_int, c_int, P, P],
    "gldm_ball_query": [P, P, c_int, c_int, c_int, c_float, c_int, P, P],
    "gldm_grouping_forward": [P, P, c_int, c_int, c_int, c_int, c_int, P, P],
    "gldm_grouping_backward": [P, P, c_int, c_int, c_int, c_int, c_int, P, P],
    "gldm_three_nn_interpolate_forward": [P, P, P, c_int, c_int, c_int, c_int, P, P, P, P],
    "gldm_three_nn_interpolate_backward": [P, P, P, c_int, c_int, c_int, c_int, P, P],
    "gldm_voxelize_fused": [P, P, c_int, c_int, c_int, c_int, P, P, P, P],
    "gldm_voxelize_fused_cl": [P, P, c_int, c_int, c_int, c_int, P, c_int, P, P],
    "gldm_pointwise_conv_f32": [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P],
    "gldm_conv3d_k3_f32": [P, P, P, c_int, c_int, c_int, c_int, P, P],
    "gldm_groupnorm_swish_f32": [P, P, P, c_int, c_int, c_int, c_int, c_float, P, P],
    "gldm_se_gate_f32": [P, P, P, c_int, c_int, c_int, P, P],
    "gldm_devox_gate_add_f32": [P, P, P, P, c_int, c_int, c_int, c_int, P, P],
    "gldm_linear_lastdim_f32": [P, P, P, c_int, c_int, c_int, P, P],
    "gldm_resnet_raw_floats": [POINTER(GldmResNetCfg)],
    "gldm_resnet_prepared_floats": [POINTER(GldmResNetCfg)],
    "gldm_resnet_prepare": [POINTER(GldmResNetCfg), P, P, P],
    "gldm_sampler_run_f32": [POINTER(GldmResNetCfg), P, P, P, c_int, c_int, c_int, P, P, c_int, c_int, P,
                             c_ulonglong, P, P, P],
    "gldm_denoiser_forward_f32": [POINTER(GldmResNetCfg), P, P, P, P, c_int, P, P],
    "gldm_denoiser_forward_f32_ftime": [POINTER(GldmResNetCfg), P, P, P, P, c_int, P, P],
    "gldm_decoder_forward_f32": [POINTER(GldmResNetCfg), P, P, c_int, P, P, c_int, c_int, P, P, P],
    "gldm_sampler_tc_pack_bytes": [POINTER(GldmResNetCfg)],
    "gldm_sampler_tc_set_profile": [P],
    "gldm_sampler_tc_set_sets": [c_int],
    "gldm_sampler_tc_set_rows": [c_int],
    "gldm_sampler_tc_prepare": [POINTER(GldmResNetCfg), P, P, P],
    "gldm_sampler_run_tc": [POINTER(GldmResNetCfg), P, P, P, P, c_int, c_int, c_int, P, P, c_int, c_int, P,
                            c_ulonglong, P, P, P],
    "gldm_time_embed_table": [POINTER(GldmResNetCfg), P, P, c_int, P, P],
    "gldm_sampler_run_tc_dev": [POINTER(GldmResNetCfg), P, P, P, P, c_int, c_int, c_int, P, P, c_int, c_int, P,
                                c_ulonglong, P, P, P],
    "gldm_denoiser_forward_tc": [POINTER(GldmResNetCfg), P, P, P, P, P, c_int, P, P],
    "gldm_denoiser_forward_tc_ftime": [POINTER(GldmResNetCfg), P, P, P, P, P, c_int, P, P],
    "gldm_gemm_tc_image_bytes": [c_longlong, c_int],
    "gldm_gemm_tc_pack_weight": [P, c_int, c_int, P, P],
    "gldm_gemm_tc_to_image": [P, c_int, c_int, c_int, P, P],
    "gldm_gemm_tc_run": [P, P, P, P, c_longlong, c_int, c_int, c_int, P, P],
    "gldm_gemm_tc_image_small_co": [P, P, P, c_longlong, c_int, c_int, c_int, P, P],
    "gldm_gemm_tc_run_proj": [P, P, P, P, c_longlong, c_int, c_int, c_int, P, P, c_int, c_int, P, P, P],
    "gldm_conv3d_tc_weight_bytes": [c_int],
    "gldm_conv3d_tc_grid_bytes": [c_int, c_int, c_int],
    "gldm_conv3d_tc_pack_weight": [P, c_int, c_int, P, P],
    "gldm_conv3d_k3_tc": [P, P, P, c_int, c_int, c_int, c_int, P, P, P],
    "gldm_decoder_forward_tc": [POINTER(GldmResNetCfg), P, P, P, c_int, P, P, c_int, c_int, P, P, P],
    "gldm_cl_pad": [P, c_int, c_int, c_int, P, P],
    "gldm_conv3d_k3_f32_cl": [P, P, P, c_int, c_int, c_int, P, c_int, P, P, P],
    "gldm_voxel_ws_bytes": [c_int, c_int, c_int],
    "gldm_block_partials_to_stats": [P, c_int, c_int, P, P],
    "gldm_conv3d_tc16_weight_bytes": [],
    "gldm_conv3d_tc16_pack_weight": [P, c_int, c_int, P, P],
    "gldm_conv3d_tc16_cl": [P, P, P, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P],
    "gldm_conv3d_tc_cl": [P, P, P, c_int, c_int, c_int, c_int, P, c_int, c_int, P, P, P],
    "gldm_gn_swish_cl": [P, c_int, c_int, P, P, P, c_int, c_int, c_int, c_float, P, P, P],
    "gldm_se_gate_sum": [P, c_int, P, P, c_int, c_int, c_int, P, P],
    "gldm_devox_cl": [P, P, c_int, c_int, P, P, c_int, c_int, c_int, c_int, P, P],
    "gldm_pose_postprocess": [P, P, P, P, c_int, P, P, P, P],
    "gldm_pose_postprocess_rows": [P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P],
    "gldm_normalize_clouds": [P, P, P, P, c_int, c_int, P, P, P, P],
    "gldm_sa_mlp_max_f32": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P],
    "gldm_se_gate_relu_f32": [P, P, P, c_int, c_int, c_int, P, P],
    "gldm_sampler_run_ex_f32": [POINTER(GldmResNetCfg), P, POINTER(GldmSamplerArgs), P],
    "gldm_sampler_run_ex_tc": [POINTER(GldmResNetCfg), P, P, POINTER(GldmSamplerArgs), P],
    "gldm_denoiser_forward_ex_f32": [POINTER(GldmResNetCfg), P, P, P, P, P, P, c_int, P, P],
    "gldm_denoiser_forward_ex_tc": [POINTER(GldmResNetCfg), P, P, P, P, P, P, P, c_int, P, P],
    "gldm_time_embed_table_f": [POINTER(GldmResNetCfg), P, P, c_int, P, P],
    "gldm_class_embed": [P, P, P, c_int, c_int, P, P],
}
_SIGS.update({"gldm_last_error": [], "gldm_version": [], "gldm_launch_count": []})
_RESTYPES = {"gldm_last_error": c_char_p, "gldm_launch_count": c_ulonglong, "gldm_voxel_ws_bytes": c_longlong, "gldm_conv3d_tc16_weight_bytes": c_longlong,
             "gldm_resnet_raw_floats": c_longlong, "gldm_sampler_tc_pack_bytes": c_longlong, "gldm_gemm_tc_image_bytes": c_longlong, "gldm_conv3d_tc_weight_bytes": c_longlong,
             "gldm_conv3d_tc_grid_bytes": c_longlong, "gldm_resnet_prepared_floats": c_longlong}

_lib = None


def exported_symbols():
    """Every symbol include/graspldm_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGS)


def lib():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(graspldm_b200 has no CPU/PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(L, name)   # AttributeError if the ABI and the header disagree
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, c_int)
        # development switches (sampler kernel variants); the defaults are chosen in csrc/sampler_tc.cu
        if os.environ.get("GLDM_TC_ROWS") is not None:
            L.gldm_sampler_tc_set_rows(int(os.environ["GLDM_TC_ROWS"]))
        _lib = L
    return _lib


class GldmError(RuntimeError):
    pass


def call(name, *args):
    """Invoke an int-returning entry point; non-zero status becomes a RuntimeError (as TORCH_CHECK does
    in the reference, R/.../functional/src/utils.hpp:7-18) instead of the reference's exit(-1)."""
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise GldmError(f"{name} failed ({rc}): {L.gldm_last_error().decode()}")


def launch_count():
    return int(lib().gldm_launch_count())
