"""ctypes binding of libepoch_b200.so (the C ABI declared in include/epoch_b200.h).

There is no CPU fallback: if the CUDA library is missing or no device is present,
loading / creating a handle fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libepoch_b200.so")

EPB_OK = 0
FIELD_NAMES = ("ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz")

# every symbol include/epoch_b200.h declares
SYMBOLS = (
    "epb_create", "epb_destroy", "epb_last_error", "epb_version", "epb_abi_info", "epb_set_stream", "epb_synchronize",
    "epb_nccl_unique_id", "epb_set_comm", "epb_upload_field", "epb_download_field", "epb_download_field_async", "epb_wait_downloads",
    "epb_upload_species", "epb_download_species", "epb_append_species", "epb_species_count", "epb_load_uniform",
    "epb_cell_counts", "epb_field_device_ptr", "epb_set_laser_source", "epb_init_boundaries",
    "epb_fields_half", "epb_push", "epb_current_finish", "epb_fields_final", "epb_sort",
    "epb_global_count", "epb_launch_count", "epb_push_kernel_ms", "epb_field_energy", "epb_step_scalars_async", "epb_wait_scalars",
    "epb_kinetic_energy", "epb_calc_moment", "epb_load_profile", "epb_redistribute", "epb_shift_window", "epb_collide", "epb_collide_pairs_test", "epb_set_boundary_temperature",
)


class Config(C.Structure):
    _fields_ = [
        ("ndims", C.c_int32), ("n", C.c_int32 * 3), ("n_global", C.c_int32 * 3), ("ng", C.c_int32),
        ("bc_field", C.c_int32 * 6), ("is_boundary", C.c_int32 * 6), ("neighbour", C.c_int32 * 27),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("n_species", C.c_int32),
        ("strict_fp", C.c_int32), ("sort_interval", C.c_int32), ("field_order", C.c_int32),
        ("maxwell_solver", C.c_int32), ("smooth_its", C.c_int32), ("smooth_comp_its", C.c_int32),
        ("smooth_strides", C.c_int32), ("hc_push", C.c_int32),
        ("dx", C.c_double * 3), ("dt", C.c_double), ("grid_min_local", C.c_double * 3),
        ("min_local", C.c_double * 3), ("max_local", C.c_double * 3),
        ("gmin", C.c_double * 3), ("gmax", C.c_double * 3),
        ("min_outer", C.c_double * 3), ("max_outer", C.c_double * 3), ("stencil", C.c_double * 15),
        ("cpml_kappa_max", C.c_double), ("cpml_a_max", C.c_double), ("cpml_sigma_max", C.c_double),
        ("cpml_thickness", C.c_int32), ("n_global_min", C.c_int32 * 3),
    ]


class SpeciesCfg(C.Structure):
    _fields_ = [
        ("charge", C.c_double), ("mass", C.c_double), ("bc_particle", C.c_int32 * 6),
        ("zero_current", C.c_int32), ("immobile", C.c_int32), ("capacity", C.c_int64),
    ]


class Decomp(C.Structure):
    """struct epb_decomp: cell_x_min(1:nprocx) ... of a tensor-product decomposition (mpi_routines.F90:317-351)"""
    _fields_ = [("nproc", C.c_int32 * 3), ("cell_min", C.POINTER(C.c_int32) * 3), ("cell_max", C.POINTER(C.c_int32) * 3)]


class Collisions(C.Structure):
    """struct epb_collisions (the collisions block of the deck, deck/deck_collision_block.F90)"""
    _fields_ = [("n_species", C.c_int32), ("coll_n_step", C.c_int32), ("use_nanbu", C.c_int32), ("reserved", C.c_int32),
                ("coulomb_log", C.c_double), ("seed", C.c_uint64), ("coll_pairs", C.POINTER(C.c_double))]


_lib = None


def load():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C epoch_b200/csrc).  epoch_b200 has no CPU fallback.")
    # libepoch_b200.so needs libnccl.so.2.  If PyTorch is importable its bundled (newer) NCCL must be
    # the one copy in the process, whichever of the two libraries gets loaded first; a Fortran host
    # simply links the system NCCL.
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec is not None and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
    except Exception:
        pass
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dp = C.c_void_p, C.c_int, C.c_int64, C.c_void_p
    L.epb_create.argtypes = [C.POINTER(Config), C.POINTER(SpeciesCfg), C.POINTER(vp)]
    L.epb_destroy.argtypes = [vp]
    L.epb_last_error.argtypes = [vp]; L.epb_last_error.restype = C.c_char_p
    L.epb_version.restype = C.c_char_p
    L.epb_abi_info.argtypes = [C.POINTER(C.c_int32)]
    info = (C.c_int32 * 4)()
    L.epb_abi_info(info)
    if info[0] != C.sizeof(Config) or info[1] != C.sizeof(SpeciesCfg):
        raise RuntimeError(f"ABI mismatch: library has sizeof(epb_config)={info[0]}, sizeof(epb_species)={info[1]}; "
                           f"binding has {C.sizeof(Config)}, {C.sizeof(SpeciesCfg)}")
    L.epb_set_stream.argtypes = [vp, vp]
    L.epb_synchronize.argtypes = [vp]
    L.epb_nccl_unique_id.argtypes = [vp]
    L.epb_set_comm.argtypes = [vp, vp]
    L.epb_upload_field.argtypes = [vp, i32, dp]
    L.epb_download_field.argtypes = [vp, i32, dp]
    L.epb_download_field_async.argtypes = [vp, i32, C.c_void_p]
    L.epb_wait_downloads.argtypes = [vp]
    L.epb_upload_species.argtypes = [vp, i32, i64, dp]
    L.epb_download_species.argtypes = [vp, i32, i64, dp]
    L.epb_append_species.argtypes = [vp, i32, i64, dp]
    L.epb_species_count.argtypes = [vp, i32, C.POINTER(i64)]
    L.epb_load_uniform.argtypes = [vp, i32, C.c_int32, C.c_double, dp, dp, C.c_uint64]
    L.epb_cell_counts.argtypes = [vp, i32, dp]
    L.epb_field_device_ptr.argtypes = [vp, i32, C.POINTER(vp)]
    L.epb_set_laser_source.argtypes = [vp, i32, dp, dp]
    for name in ("epb_init_boundaries", "epb_fields_half", "epb_push", "epb_current_finish",
                 "epb_fields_final", "epb_sort"):
        getattr(L, name).argtypes = [vp]
    L.epb_global_count.argtypes = [vp, i32, C.POINTER(i64)]
    L.epb_field_energy.argtypes = [vp, dp]
    L.epb_step_scalars_async.argtypes = [vp, dp, C.POINTER(C.c_int64)]
    L.epb_wait_scalars.argtypes = [vp, C.c_int64]
    L.epb_kinetic_energy.argtypes = [vp, i32, C.POINTER(C.c_double)]
    L.epb_calc_moment.argtypes = [vp, i32, i32, dp]
    L.epb_set_boundary_temperature.argtypes = [vp, i32, i32, dp]
    L.epb_collide.argtypes = [vp, C.POINTER(Collisions)]
    L.epb_collide_pairs_test.argtypes = [C.c_int, dp, dp, dp, dp, dp, dp, dp]
    L.epb_redistribute.argtypes = [vp, C.POINTER(Decomp), C.POINTER(Decomp), C.POINTER(Config), C.POINTER(SpeciesCfg),
                                   C.POINTER(vp)]
    L.epb_shift_window.argtypes = [vp, C.POINTER(Decomp), C.POINTER(Config), C.POINTER(SpeciesCfg), C.c_double,
                                   C.POINTER(vp)]
    L.epb_load_profile.argtypes = [vp, i32, dp]
    L.epb_launch_count.argtypes = [vp]; L.epb_launch_count.restype = i64
    L.epb_push_kernel_ms.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64), i32]
    _lib = L
    return L
