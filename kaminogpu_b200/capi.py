"""ctypes binding of the kamino_b200 C ABI (include/kamino_b200.h).

Loads the in-tree ``libkamino_b200.so`` (built by ``kaminogpu_b200/Makefile`` /
``__graft_entry__.build()``). There is no CPU fallback: importing works without a GPU
(so that symbol checks can run anywhere), but creating a context raises ``KaminoError``
when no sm_100 device is present, and a missing library raises ``ImportError``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KAMINO_B200_LIB: development aid of this binding (A/B runs of alternative builds of the library), not read by the library
LIB_PATH = os.environ.get("KAMINO_B200_LIB") or os.path.join(_HERE, "libkamino_b200.so")

# field ids (include/kamino_b200.h)
VEL_PHI, VEL_THETA, DENSITY, PRESSURE = 0, 1, 2, 3
SAMPLE_VPHI, SAMPLE_VTHETA, SAMPLE_CENTERED = 0, 1, 2

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int32_p = ctypes.POINTER(ctypes.c_int32)

# name -> (restype, argtypes); every symbol the header declares
SIGNATURES = {
    "kamino_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
                                     ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_long]),
    "kamino_alloc_particles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_long]),
    "kamino_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "kamino_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "kamino_get_shape": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_long)]),
    "kamino_upload_field": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "kamino_download_field": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "kamino_upload_particles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kamino_download_particles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kamino_download_field_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "kamino_download_particles_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kamino_upload_field_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "kamino_upload_particles_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kamino_field_device_ptr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    "kamino_particles_device_ptr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                   ctypes.POINTER(ctypes.c_void_p)]),
    "kamino_advect": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_geometric": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_project": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_band_advect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "kamino_band_geometric": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "kamino_band_divergence_fft": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "kamino_band_tridiagonal": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "kamino_band_inverse_fft_gradient": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "kamino_band_solver_prepare": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "kamino_band_local_solve": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_tridiagonal_coefficients": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_float_p, c_float_p]),
    "kamino_spectrum_device_ptr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "kamino_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "kamino_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_run_frames": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "kamino_phase_times": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_int]),
    "kamino_launches_per_step": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_profile_steps": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_float_p]),
    "kamino_init_velocity_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "kamino_init_velocity_host_rows": (ctypes.c_int, [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                      ctypes.c_void_p, ctypes.c_void_p]),
    "kamino_dist_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_dist_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                          ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "kamino_dist_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_dist_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "kamino_dist_shape": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(ctypes.c_size_t)]),
    "kamino_dist_upload": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kamino_dist_download": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kamino_dist_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "kamino_dist_group_step": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int]),
    "kamino_dist_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_dist_transport": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_char_p)]),
    "kamino_dist_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "kamino_dist_comm_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                              ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_long),
                                              ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "kamino_init_velocity_device": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_dist_init_velocity_device": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_seed_particles_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float, ctypes.c_ulonglong]),
    "kamino_particle_count": (ctypes.c_long, [ctypes.c_int, ctypes.c_float]),
    "kamino_seed_particles_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_float, ctypes.c_void_p]),
    "kamino_debug_locate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "kamino_debug_project_cr": (ctypes.c_int, [ctypes.c_void_p]),
    "kamino_version": (ctypes.c_char_p, []),
    "kamino_host_alloc": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]),
    "kamino_host_free": (ctypes.c_int, [ctypes.c_void_p]),
}


class KaminoError(RuntimeError):
    """A non-zero return code from the C ABI."""

    def __init__(self, code, message):
        super().__init__("kamino_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def load():
    """Load libkamino_b200.so and attach the signatures. Raises ImportError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "kaminogpu_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C kaminogpu_b200` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(code, ctx=None):
    if code != 0:
        msg = load().kamino_last_error(ctx)
        raise KaminoError(code, msg.decode() if msg else "")
