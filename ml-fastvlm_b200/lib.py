"""ctypes binding of libfastvithd_b200.so (C ABI: include/fastvithd_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libfastvithd_b200.so"
_lib = None

F32, F16, BF16 = 0, 1, 2


class FvhdError(RuntimeError):
    """Raised for every non-zero status of the C library (message from fvhd_last_error)."""


class FvhdConfig(C.Structure):
    _fields_ = [("image_size", C.c_int), ("projector_hidden", C.c_int), ("projector_depth", C.c_int), ("max_batch", C.c_int)]


class FvhdLlmConfig(C.Structure):
    _fields_ = [("hidden", C.c_int), ("layers", C.c_int), ("heads", C.c_int), ("kv_heads", C.c_int), ("head_dim", C.c_int),
                ("intermediate", C.c_int), ("vocab", C.c_int), ("max_seq", C.c_int), ("rope_theta", C.c_float), ("rms_eps", C.c_float)]


class FvhdTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int), ("numel", C.c_int64)]


# every symbol include/fastvithd_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "fvhd_api_version": (C.c_int, []),
    "fvhd_create": (C.c_int, [C.POINTER(FvhdConfig), C.POINTER(C.c_void_p)]),
    "fvhd_destroy": (C.c_int, [C.c_void_p]),
    "fvhd_last_error": (C.c_char_p, [C.c_void_p]),
    "fvhd_num_weights": (C.c_int, [C.c_void_p]),
    "fvhd_weight_spec": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "fvhd_load_weights": (C.c_int, [C.c_void_p, C.POINTER(FvhdTensor), C.c_int]),
    "fvhd_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "fvhd_set_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "fvhd_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fvhd_forward_strided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong]),
    "fvhd_forward_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_int]),
    "fvhd_forward_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fvhd_encode_images_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fvhd_num_tokens": (C.c_int, [C.c_void_p]),
    "fvhd_out_dim": (C.c_int, [C.c_void_p]),
    "fvhd_num_units": (C.c_int, [C.c_void_p]),
    "fvhd_unit_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "fvhd_run_units": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fvhd_profile_units": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int]),
    "fvhd_num_steps": (C.c_int, [C.c_void_p, C.c_int]),
    "fvhd_step_info": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "fvhd_profile_steps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int]),
    "fvhd_launches_per_forward": (C.c_int, [C.c_void_p, C.c_int]),
    "fvhd_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "fvhd_preprocess_tiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int]),
    "fvhd_resample_coeffs": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]),
    "fvhd_debug_gemm_trace": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "fvhd_debug_mixer_trace": (C.c_int, [C.c_void_p]),
    "fvhd_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_int, C.c_int, C.c_int, C.c_int]),
    "fvhd_convffn_half": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_int]),
    "fvhd_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "fvhd_mixer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_int, C.c_int, C.c_int, C.c_int]),
    "fvhd_convffn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int, C.c_int, C.c_void_p]),
    "fvhd_llm_load": (C.c_int, [C.c_void_p, C.POINTER(FvhdLlmConfig), C.POINTER(C.c_void_p), C.c_int]),
    "fvhd_llm_input": (C.c_void_p, [C.c_void_p]),
    "fvhd_llm_prefill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "fvhd_llm_kv_cache": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "fvhd_llm_launches": (C.c_int, [C.c_void_p, C.c_int]),
}


def library_path():
    return os.path.join(_HERE, _LIB_NAME)


def load_library():
    """Load the CUDA library.  Fails loudly when it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise FvhdError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.fvhd_api_version() != 1:
        raise FvhdError(f"{path}: API version {lib.fvhd_api_version()} != 1")
    _lib = lib
    return lib


def check(rc, handle=None):
    if rc != 0:
        msg = load_library().fvhd_last_error(handle)
        raise FvhdError(f"libfastvithd_b200 status {rc}: {msg.decode() if msg else '?'}")
