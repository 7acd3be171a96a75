"""ctypes binding of include/bayesvg_b200.h (the C ABI an R package would reach through .Call).

There is no CPU fallback: if libbayesvg_b200.so is missing this module raises at import of the
library handle, and every compute entry point fails with BVG_ERR_CUDA when no device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libbayesvg_b200.so")

ABI_VERSION = 2
ALG_MEANFIELD, ALG_FULLRANK, ALG_PATHFINDER = 0, 1, 2
GAUSSIAN, NB = 0, 1
KERNELS = {"exp_quad": 0, "matern": 1, "periodic": 2}
PREC_FP64, PREC_FAST = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2

EXPORTS = [
    "bvg_version", "bvg_last_error", "bvg_release_cached_memory", "bvg_device_count", "bvg_config_default", "bvg_basis_build", "bvg_kmeans", "bvg_variogram_lscale", "bvg_gmm_log_prob", "bvg_gmm_fit", "bvg_gmm_open", "bvg_gmm_close", "bvg_gmm_eval", "bvg_gmm_posterior",
    "bvg_fit_create", "bvg_fit_destroy", "bvg_fit_set_phi", "bvg_fit_set_y_f64", "bvg_fit_set_y_i32",
    "bvg_fit_dim", "bvg_fit_shape", "bvg_fit_set_gene_depths", "bvg_comm_unique_id", "bvg_fit_comm_init", "bvg_fit_peer_handle", "bvg_fit_peer_init", "bvg_log_prob", "bvg_fit_run", "bvg_fit_mle",
    "bvg_fit_advi", "bvg_fit_summarize", "bvg_fit_set_variational", "bvg_fit_get_variational", "bvg_fit_get_cholesky", "bvg_fit_set_cholesky", "bvg_fit_pathfinder_single",
    "bvg_fit_advi_steps", "bvg_fit_elbo", "bvg_fit_time_lik_kernel", "bvg_fit_flush_l2", "bvg_fit_get_summary",
    "bvg_fit_get_draws", "bvg_fit_posterior_ncol", "bvg_fit_get_posterior", "bvg_fit_get_trace", "bvg_fit_get_stats", "bvg_debug_stage_times", "bvg_debug_gene_trace", "bvg_debug_time_gene",
    "bvg_debug_tc_trace", "bvg_debug_persist_trace",
]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("likelihood", C.c_int32), ("precision", C.c_int32), ("engine", C.c_int32),
        ("device", C.c_int32), ("n_iter", C.c_int32), ("mle_init", C.c_int32), ("mle_iter", C.c_int32),
        ("n_draws", C.c_int32), ("elbo_samples", C.c_int32), ("eval_elbo", C.c_int32), ("adapt_iter", C.c_int32),
        ("lbfgs_history", C.c_int32), ("verbose", C.c_int32), ("tol_rel_obj", C.c_double), ("eta", C.c_double),
        ("seed", C.c_uint64), ("rank", C.c_int32), ("world_size", C.c_int32), ("gene_offset", C.c_int64),
        ("G_total", C.c_int64), ("elbo_suffstat", C.c_int32), ("depth_adjust", C.c_int32),
        ("algorithm", C.c_int32), ("pf_num_paths", C.c_int32), ("pf_max_lbfgs_iters", C.c_int32),
        ("pf_single_draws", C.c_int32), ("pf_init_radius", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("eta", C.c_double), ("advi_iters", C.c_int32), ("grad_evals", C.c_int32), ("elbo_evals", C.c_int32),
        ("converged", C.c_int32), ("mle_iters", C.c_int32), ("mle_fevals", C.c_int32), ("mle_term", C.c_int32),
        ("mle_lp", C.c_double), ("elbo", C.c_double), ("sec_h2d", C.c_double), ("sec_mle", C.c_double),
        ("sec_adapt", C.c_double), ("sec_sga", C.c_double), ("sec_draws", C.c_double), ("sec_total", C.c_double),
        ("kernel_launches", C.c_int64), ("engine_used", C.c_int32), ("mle_fp64_from", C.c_int32),
    ]


class BvgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bayesvg_b200 error {code}: {msg}")
        self.code = code


_lib = None
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_vp = C.c_void_p


def lib():
    """Load the shared library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C bayesvg_b200/csrc). bayesvg_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
    L.bvg_last_error.restype = C.c_char_p
    L.bvg_config_default.argtypes = [C.POINTER(Config)]
    L.bvg_config_default.restype = None
    L.bvg_basis_build.argtypes = [C.c_int, _dp, C.c_int64, _dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, _dp]
    L.bvg_fit_create.argtypes = [C.POINTER(Config), C.POINTER(_vp)]
    L.bvg_fit_destroy.argtypes = [_vp]
    L.bvg_fit_destroy.restype = None
    L.bvg_fit_set_phi.argtypes = [_vp, _dp, C.c_int64, C.c_int]
    L.bvg_fit_set_y_f64.argtypes = [_vp, _dp, C.c_int64, C.c_int64]
    L.bvg_fit_set_y_i32.argtypes = [_vp, C.POINTER(C.c_int32), C.c_int64, C.c_int64]
    L.bvg_fit_dim.argtypes = [_vp, C.POINTER(C.c_int64)]
    L.bvg_fit_shape.argtypes = [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    L.bvg_fit_set_gene_depths.argtypes = [_vp, _dp, C.c_int64]
    L.bvg_comm_unique_id.argtypes = [_vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.bvg_fit_comm_init.argtypes = [_vp, _vp, C.c_size_t]
    L.bvg_fit_peer_handle.argtypes = [_vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.bvg_fit_peer_init.argtypes = [_vp, _vp, C.c_size_t]
    L.bvg_log_prob.argtypes = [_vp, _dp, C.c_int, C.c_int, _dp, _dp]
    L.bvg_fit_run.argtypes = [_vp]
    L.bvg_fit_mle.argtypes = [_vp, _dp, _dp]
    L.bvg_fit_advi.argtypes = [_vp, _dp]
    L.bvg_fit_summarize.argtypes = [_vp]
    L.bvg_fit_set_variational.argtypes = [_vp, _dp, _dp]
    L.bvg_fit_get_variational.argtypes = [_vp, _dp, _dp]
    L.bvg_fit_advi_steps.argtypes = [_vp, C.c_int, C.c_double, C.c_int, _fp]
    L.bvg_fit_elbo.argtypes = [_vp, C.c_int, _dp, _fp]
    L.bvg_fit_time_lik_kernel.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _fp]
    L.bvg_fit_flush_l2.argtypes = [_vp]
    L.bvg_debug_stage_times.argtypes = [_vp, C.c_int, _fp]
    L.bvg_debug_time_gene.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]
    L.bvg_debug_tc_trace.argtypes = [_vp, C.POINTER(C.c_longlong), C.c_int]
    L.bvg_debug_gene_trace.argtypes = [_vp, C.POINTER(C.c_longlong), C.c_int]
    L.bvg_debug_persist_trace.argtypes = [_vp, C.c_int, C.POINTER(C.c_longlong), C.c_int]
    L.bvg_fit_get_summary.argtypes = [_vp, _dp]
    L.bvg_fit_get_draws.argtypes = [_vp, _dp]
    L.bvg_gmm_log_prob.argtypes = [C.c_int, _dp, C.c_int64, C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.POINTER(C.c_double), _dp]
    L.bvg_gmm_fit.argtypes = [C.c_int, _dp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64, _dp, _dp, _dp, _dp, _dp, _dp]
    L.bvg_release_cached_memory.argtypes = []
    L.bvg_release_cached_memory.restype = None
    L.bvg_gmm_open.argtypes = [C.c_int, _dp, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.bvg_gmm_close.argtypes = [C.c_void_p]
    L.bvg_gmm_close.restype = None
    L.bvg_gmm_eval.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp]
    L.bvg_gmm_posterior.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp, C.POINTER(C.c_double)]
    L.bvg_variogram_lscale.argtypes = [C.c_int, _dp, C.c_int64, _dp, C.c_int64, C.c_double, _dp, _dp]
    L.bvg_fit_pathfinder_single.argtypes = [_vp, C.c_int, _dp, _dp, _dp, _dp]
    L.bvg_fit_get_cholesky.argtypes = [_vp, _dp]
    L.bvg_fit_set_cholesky.argtypes = [_vp, _dp]
    L.bvg_kmeans.argtypes = [C.c_int, _dp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64, _dp, _dp, C.POINTER(C.c_int32)]
    L.bvg_fit_posterior_ncol.argtypes = [_vp, C.POINTER(C.c_int64)]
    L.bvg_fit_get_posterior.argtypes = [_vp, C.c_int, _dp]
    L.bvg_fit_get_trace.argtypes = [_vp, _dp, C.c_int, C.POINTER(C.c_int)]
    L.bvg_fit_get_stats.argtypes = [_vp, C.POINTER(Stats)]
    if L.bvg_version() != ABI_VERSION:
        raise ImportError(f"libbayesvg_b200.so ABI {L.bvg_version()} != binding {ABI_VERSION}")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise BvgError(rc, lib().bvg_last_error().decode("utf-8", "replace"))
