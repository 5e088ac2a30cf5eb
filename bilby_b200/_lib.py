"""ctypes binding of the C ABI (include/bilby_b200.h).  Fails loudly when the CUDA library or a
CUDA device is missing - there is no CPU path."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BILBY_B200_LIB") or os.path.join(_HERE, "_lib", "libbilby_b200.so")

c_double_p = ctypes.c_void_p
_lib = None


class BilbyB200Error(RuntimeError):
    pass


def load():
    """Load (building first if the shared object is absent) and declare every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build
        build.build()
    lib = ctypes.CDLL(LIB_PATH)
    vp, i, d, lng = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_long
    sigs = {
        "bb_last_error": (ctypes.c_char_p, []),
        "bb_abi_version": (i, []),
        "bb_create": (i, [i, ctypes.POINTER(vp)]),
        "bb_destroy": (None, [vp]),
        "bb_set_network": (i, [vp, i, i, d, d, d, vp, vp, vp, vp, vp]),
        "bb_set_waveform": (i, [vp, i, d, d, d]),
        "bb_set_marginalization": (i, [vp, i, d, vp, i, vp, i, vp, d, d, d, d, d, d, i]),
        "bb_log_likelihood_ratio_device": (i, [vp, vp, lng, vp, vp]),
        "bb_log_likelihood_ratio_host": (i, [vp, vp, lng, vp]),
        "bb_inner_products_device": (i, [vp, vp, lng, vp, vp]),
        "bb_set_calibration": (i, [vp, i, vp, vp, vp]),
        "bb_log_likelihood_ratio_cal_device": (i, [vp, vp, vp, lng, vp, vp]),
        "bb_log_likelihood_ratio_cal_host": (i, [vp, vp, vp, lng, vp]),
        "bb_inner_products_cal_device": (i, [vp, vp, vp, lng, vp, vp]),
        "bb_likelihood_from_inner_products_device": (i, [vp, vp, vp, lng, vp, vp]),
        "bb_set_frequency_shard": (i, [vp, i, i]),
        "bb_set_reference_frame": (i, [vp, vp, vp]),
        "bb_sky_frame_parameters_device": (i, [vp, vp, lng, vp, vp]),
        "bb_frequency_domain_strain_device": (i, [vp, vp, lng, vp, vp]),
        "bb_frequency_sequence_strain_device": (i, [vp, vp, lng, vp, i, d, vp, vp]),
        "bb_set_relative_binning": (i, [vp, i, vp, vp, vp, vp, vp]),
        "bb_set_roq": (i, [vp, i, vp, i, vp, i, lng, d, vp, vp, i, d, d, d]),
        "bb_set_multiband": (i, [vp, i, vp, vp, vp]),
        "bb_set_multiband_time_marginalization": (i, [vp, lng, vp, d, d]),
        "bb_set_multiband_ifft_fft": (i, [vp, i, vp, vp, vp, vp, vp, vp, vp]),
        "bb_fft_device": (i, [vp, vp, vp, lng, i, vp]),
        "bb_exchange_create": (i, [vp, i, i, lng, vp]),
        "bb_exchange_connect": (i, [vp, vp]),
        "bb_log_likelihood_ratio_sharded_device": (i, [vp, vp, lng, vp, vp]),
        "bb_exchange_status": (i, [vp, ctypes.POINTER(ctypes.c_int)]),
        "bb_exchange_destroy": (i, [vp]),
        "bb_detector_response_device": (i, [vp, vp, lng, vp, vp]),
        "bb_build_distance_table": (i, [vp, vp, i, vp, i, vp, vp, i, d, i, vp]),
        "bb_antenna_response_device": (i, [vp, vp, lng, vp, vp]),
        "bb_ln_i0_device": (i, [vp, vp, lng, vp, vp]),
        "bb_project_polarizations_device": (i, [vp, i, vp, vp, vp, vp, vp]),
        "bb_noise_weighted_inner_product_device": (i, [vp, i, vp, vp, vp, vp]),
        "bb_profile_enable": (i, [vp, i]),
        "bb_profile_read": (i, [vp, ctypes.POINTER(d), ctypes.POINTER(lng)]),
        "bb_fp64_peak": (i, [vp, ctypes.POINTER(d)]),
        "bb_fp64_tensor_peak": (i, [vp, ctypes.POINTER(d)]),
        "bb_build_roq_quadratic_weights": (i, [i, i, i, vp, i, vp, d, vp]),
        "bb_build_relbin_summary_data": (i, [vp, i, vp, vp, vp, vp]),
        "bb_launch_count": (lng, [vp]),
        "bb_set_reconstruction_grid": (i, [vp, vp, vp, i]),
        "bb_set_calibration_marginalization": (i, [vp, i, vp]),
        "bb_build_roq_linear_weights": (i, [i, i, i, vp, i, vp, vp, lng, lng, i, d, vp]),
        "bb_reconstruct_marginalized_device": (i, [vp, vp, vp, lng, vp, vp, vp]),
        "bb_contract_device": (i, [vp, i, i, i, i, i, lng, lng, i, lng, lng, lng, d, vp, lng, vp, lng, i, vp, lng, vp]),
        "bb_set_sampling_priors": (i, [vp, i, vp, i, vp, vp, i]),
        "bb_rows_from_unit_cube_device": (i, [vp, vp, lng, vp, vp, vp]),
        "bb_rows_from_theta_device": (i, [vp, vp, lng, vp, vp]),
        "bb_math_device": (i, [vp, i, vp, lng, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = (
    "bb_last_error", "bb_abi_version", "bb_create", "bb_destroy", "bb_set_network", "bb_set_waveform",
    "bb_set_marginalization", "bb_log_likelihood_ratio_device", "bb_log_likelihood_ratio_host",
    "bb_inner_products_device", "bb_set_calibration", "bb_log_likelihood_ratio_cal_device",
    "bb_log_likelihood_ratio_cal_host", "bb_inner_products_cal_device", "bb_likelihood_from_inner_products_device", "bb_set_frequency_shard", "bb_set_reference_frame", "bb_sky_frame_parameters_device",
    "bb_frequency_domain_strain_device", "bb_frequency_sequence_strain_device", "bb_set_relative_binning",
    "bb_set_roq", "bb_detector_response_device", "bb_build_distance_table",
    "bb_antenna_response_device", "bb_ln_i0_device", "bb_project_polarizations_device",
    "bb_noise_weighted_inner_product_device", "bb_profile_enable", "bb_profile_read", "bb_fp64_peak",
    "bb_launch_count", "bb_set_reconstruction_grid", "bb_reconstruct_marginalized_device",
    "bb_set_calibration_marginalization", "bb_build_roq_linear_weights", "bb_set_multiband",
    "bb_exchange_create", "bb_exchange_connect", "bb_log_likelihood_ratio_sharded_device", "bb_exchange_status",
    "bb_exchange_destroy", "bb_contract_device", "bb_fp64_tensor_peak",
    "bb_build_roq_quadratic_weights", "bb_build_relbin_summary_data", "bb_set_multiband_time_marginalization",
    "bb_set_multiband_ifft_fft", "bb_fft_device", "bb_set_sampling_priors", "bb_rows_from_unit_cube_device",
    "bb_rows_from_theta_device", "bb_math_device")


_torch_ops = None


def torch_ops():
    """torch.ops.bilby_b200: the TORCH_LIBRARY shim over the C ABI (csrc/bb_torch.cpp).  Built by bilby_b200.build."""
    global _torch_ops
    if _torch_ops is None:
        import torch
        load()                                   # the shim links against libbilby_b200.so
        # always the in-tree shim: its DT_NEEDED names the soname libbilby_b200.so, which the copy ctypes has already
        # loaded satisfies (also an experiment build selected with BILBY_B200_LIB)
        path = os.path.join(_HERE, "_lib", "libbilby_b200_torch.so")
        if not os.path.exists(path):
            from . import build
            build.build_torch_shim()
        torch.ops.load_library(path)
        _torch_ops = torch.ops.bilby_b200
    return _torch_ops


def check(rc):
    if rc != 0:
        raise BilbyB200Error(load().bb_last_error().decode())


class Handle:
    """Owns one bb_handle (device tiles + scratch) on one CUDA device."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise BilbyB200Error("bilby_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        ptr = ctypes.c_void_p()
        check(self.lib.bb_create(self.device, ctypes.byref(ptr)))
        self.ptr = ptr

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.lib.bb_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass
