"""ctypes binding of include/hot_b200.h (the reference-side binding a maintainer would add is in INTEGRATION.md)."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhot_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "hot_b200.h")


class LibraryMissing(RuntimeError):
    pass


_lib = None
# hot_transport (include/hot_b200.h): the caller's collectives for a partitioned object
T_ALL_REDUCE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_int)
T_ALL_GATHER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long)
T_NEIGHBOR_EXCHANGE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_long))


class Transport(C.Structure):
    _fields_ = [("user", C.c_void_p), ("all_reduce", T_ALL_REDUCE), ("all_gather", T_ALL_GATHER), ("neighbor_exchange", T_NEIGHBOR_EXCHANGE)]



HOT_LOG_CAP = 256


class SolverOptions(C.Structure):
    """hot_solver_options of include/hot_b200.h (field names = the reference's command-line flags)"""
    _fields_ = [(n, C.c_int) for n in ("lsolver", "matfree", "project", "bcproject", "linesearch", "usecn", "adaptive_h", "mg_level",
                                       "mg_times", "mg_scale", "smoother", "coarse_solver", "Ainv", "max_newton_iterations",
                                       "max_lbfgs_iterations", "max_cg_iterations")] + [("cneps", C.c_double), ("topomega", C.c_double)]


class SolveLog(C.Structure):
    """hot_solve_log of include/hot_b200.h"""
    _fields_ = [(n, C.c_int) for n in ("iterations", "converged", "n_log", "matrix_builds", "total_linear_iterations",
                                       "total_linesearch_probes")] + [
        ("tolerance", C.c_double), ("residual_norm", C.c_double * HOT_LOG_CAP), ("scaled_norm", C.c_double * HOT_LOG_CAP),
        ("energy", C.c_double * HOT_LOG_CAP), ("linear_iterations", C.c_int * HOT_LOG_CAP)]

    def as_dict(self):
        n = self.n_log
        return dict(iterations=self.iterations, converged=bool(self.converged), matrix_builds=self.matrix_builds,
                    total_linear_iterations=self.total_linear_iterations, total_linesearch_probes=self.total_linesearch_probes,
                    tolerance=self.tolerance, residual_norm=list(self.residual_norm[:n]), scaled_norm=list(self.scaled_norm[:n]),
                    energy=list(self.energy[:n]), linear_iterations=list(self.linear_iterations[:n]))

_c_double_p = C.POINTER(C.c_double)
_c_int_p = C.POINTER(C.c_int)
_c_u64_p = C.POINTER(C.c_ulonglong)
_c_i64_p = C.POINTER(C.c_longlong)


def declared_symbols(header=HEADER_PATH):
    """Every function name declared in the C header (used by the ABI test)."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hot_[a-z0-9_]+)\s*\(", text)))


def load_library(path=LIB_PATH):
    """Load the CUDA library; raise loudly (no CPU fallback exists) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise LibraryMissing(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(hot_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    vp = C.c_void_p
    sig = {
        "hot_create": (vp, [C.c_double, C.c_double, C.c_double, C.c_int]),
        "hot_destroy": (None, [vp]),
        "hot_last_error": (C.c_char_p, [vp]),
        "hot_set_stream": (C.c_int, [vp, vp]),
        "hot_synchronize": (C.c_int, [vp]),
        "hot_launch_count": (C.c_longlong, [vp]),
        "hot_timing": (C.c_int, [vp, C.c_int]),
        "hot_get_timings": (C.c_int, [vp, C.c_int, _c_double_p, _c_i64_p]),
        "hot_timing_name": (C.c_char_p, [C.c_int]),
        "hot_linear_offset": (C.c_int, [vp, C.c_long, _c_int_p, _c_u64_p]),
        "hot_linear_to_coord": (C.c_int, [vp, C.c_long, _c_u64_p, _c_int_p]),
        "hot_packed_add": (C.c_int, [vp, C.c_long, _c_u64_p, _c_u64_p, _c_u64_p]),
        "hot_linear_offset_f32": (C.c_int, [vp, C.c_long, _c_int_p, _c_u64_p]),
        "hot_linear_to_coord_f32": (C.c_int, [vp, C.c_long, _c_u64_p, _c_int_p]),
        "hot_packed_add_f32": (C.c_int, [vp, C.c_long, _c_u64_p, _c_u64_p, _c_u64_p]),
        "hot_set_particles": (C.c_int, [vp, C.c_long] + [vp] * 8),
        "hot_get_particles": (C.c_int, [vp] + [vp] * 5),
        "hot_upload_state_async": (C.c_int, [vp] + [vp] * 4),
        "hot_commit_state": (C.c_int, [vp]),
        "hot_download_state_async": (C.c_int, [vp] + [vp] * 4),
        "hot_wait_download": (C.c_int, [vp]),
        "hot_num_particles": (C.c_long, [vp]),
        "hot_sort_and_activate": (C.c_int, [vp]),
        "hot_num_groups": (C.c_long, [vp]),
        "hot_num_pages": (C.c_long, [vp]),
        "hot_get_sort": (C.c_int, [vp, vp, vp, vp]),
        "hot_get_groups": (C.c_int, [vp, vp, vp, vp]),
        "hot_get_pages": (C.c_int, [vp, vp]),
        "hot_p2g": (C.c_int, [vp, _c_int_p]),
        "hot_num_nodes": (C.c_int, [vp]),
        "hot_get_grid": (C.c_int, [vp, vp, vp, vp]),
        "hot_get_id2coord": (C.c_int, [vp, vp]),
        "hot_get_mass_matrix": (C.c_int, [vp, vp]),
        "hot_set_dv": (C.c_int, [vp, vp]),
        "hot_g2p": (C.c_int, [vp, C.c_double, _c_int_p]),
        "hot_set_plasticity": (C.c_int, [vp, C.c_int, vp]),
        "hot_apply_plasticity": (C.c_int, [vp]),
        "hot_get_plastic_state": (C.c_int, [vp, vp, vp, vp]),
        "hot_set_plastic_state": (C.c_int, [vp, vp]),
        "hot_comm_unique_id": (C.c_int, [vp]),
        "hot_comm_init_nccl": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "hot_set_partition": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "hot_share_tables": (C.c_int, [C.c_int, C.c_int, C.c_int] + [vp] * 3 + [vp] * 11),
        "hot_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_long]),
        "hot_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_long]),
        "hot_get_partition": (C.c_int, [vp, C.POINTER(C.c_long)]),
        "hot_get_transport": (C.c_int, [vp]),
        "hot_set_ghost_ring": (C.c_int, [vp, C.c_int]),
        "hot_set_constitutive_model": (C.c_int, [vp, C.c_int]),
        "hot_halo_pages": (C.c_int, [C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(C.c_int), vp]),
        "hot_page_authority": (C.c_int, [C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]),
        "hot_set_dt_gravity": (C.c_int, [vp, C.c_double, vp]),
        "hot_set_project": (C.c_int, [vp, C.c_int]),
        "hot_set_bc": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]),
        "hot_get_dv": (C.c_int, [vp, vp]),
        "hot_set_colliders": (C.c_int, [vp, C.c_int, vp]),
        "hot_build_bc": (C.c_int, [vp, C.c_int, _c_int_p]),
        "hot_get_bc": (C.c_int, [vp, _c_int_p, vp, vp, vp, vp, vp]),
        "hot_corotated_eval": (C.c_int, [vp, C.c_long, vp, C.c_double, C.c_double, C.c_int] + [vp] * 8),
        "hot_strain_energy": (C.c_int, [vp, vp]),
        "hot_get_strain_backup": (C.c_int, [vp, vp]),
        "hot_backup_strain": (C.c_int, [vp]),
        "hot_restore_strain": (C.c_int, [vp]),
        "hot_update_state": (C.c_int, [vp, vp, _c_double_p]),
        "hot_get_stress": (C.c_int, [vp, vp, vp]),
        "hot_compute_residual": (C.c_int, [vp, vp]),
        "hot_project": (C.c_int, [vp, vp]),
        "hot_hessian_apply_mf": (C.c_int, [vp, vp, vp]),
        "hot_add_scaled_forces": (C.c_int, [vp, C.c_double, vp]),
        "hot_add_scaled_force_differentials": (C.c_int, [vp, C.c_double, vp, vp]),
        "hot_eval_cn_tolerance": (C.c_int, [vp, C.c_double, C.c_double, vp]),
        "hot_build_matrix": (C.c_int, [vp, C.c_int]),
        "hot_get_matrix": (C.c_int, [vp, vp, vp]),
        "hot_build_diagonal": (C.c_int, [vp, C.c_int, vp]),
        "hot_build_mg": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
        "hot_mg_levels": (C.c_int, [vp]),
        "hot_estimate_2norm": (C.c_int, [vp, C.c_int, C.POINTER(C.c_double)]),
        "hot_get_level_dofs": (C.c_int, [vp, _c_int_p]),
        "hot_level_nnz_blocks": (C.c_int, [vp, C.c_int, _c_i64_p]),
        "hot_get_level_coords": (C.c_int, [vp, C.c_int, vp]),
        "hot_get_level_matrix": (C.c_int, [vp, C.c_int, C.c_int, _c_int_p, vp, vp]),
        "hot_get_level_diagonal": (C.c_int, [vp, C.c_int, vp, vp]),
        "hot_get_gs_schedule": (C.c_int, [vp, C.c_int, _c_int_p, _c_int_p, vp, vp]),
        "hot_spmv": (C.c_int, [vp, C.c_int, vp, vp]),
        "hot_restrict": (C.c_int, [vp, C.c_int, vp, vp]),
        "hot_prolong": (C.c_int, [vp, C.c_int, vp, vp]),
        "hot_smooth": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_double, vp]),
        "hot_vcycle": (C.c_int, [vp, vp, vp]),
        "hot_vcycle_timing": (C.c_int, [vp, vp, _c_int_p]),
        "hot_vcycle_bench": (C.c_int, [vp, C.c_int, _c_double_p]),
        "hot_op_bench": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, _c_double_p]),
        "hot_default_options": (None, [C.POINTER(SolverOptions)]),
        "hot_pcg": (C.c_int, [vp, vp, vp, C.c_double, C.c_int, C.c_int, C.c_int, _c_int_p]),
        "hot_get_dv0": (C.c_int, [vp, vp]),
        "hot_backward_euler_step": (C.c_int, [vp, C.POINTER(SolverOptions), C.POINTER(SolveLog)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._hot_signatures = sig
    _lib = lib
    return lib
