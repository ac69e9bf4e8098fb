"""ctypes loader for swiftest_b200/lib/libswiftest_cuda.so (C ABI: include/swiftest_cuda.h).

There is no CPU fallback: a missing library or a missing sm_100 GPU raises immediately.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libswiftest_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "swiftest_cuda.h")

_lib = None


class SwcuError(RuntimeError):
    pass


def declared_symbols(header=HEADER_PATH):
    """Every function name include/swiftest_cuda.h declares."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(swcu_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SwcuError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C swiftest_b200/csrc). swiftest_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    p, i32, i64, u64, d = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    pp = C.POINTER(C.c_void_p)
    sig = {
        "swcu_create": [C.c_int, pp],
        "swcu_destroy": [p],
        "swcu_set_stream": [p, p],
        "swcu_synchronize": [p],
        "swcu_device_info": [p, p, p, p],
        "swcu_kick_getacch_int_all_flat_pl": [p, i32, i64, p, p, p, p, p],
        "swcu_kick_getacch_int_all_tri_pl": [p, i32, i32, p, p, p, p],
        "swcu_kick_getacch_int_all_tp": [p, i32, i32, p, p, p, p, p],
        "swcu_symba_kick_subtract_encounters": [p, i32, i64, p, p, p, p, p, p],
        "swcu_drift_all": [p, i32, p, p, p, d, i32, d, p, p],
        "swcu_encounter_check_all_sort_and_sweep_plpl": [p, i32, p, p, p, d, p],
        "swcu_encounter_check_all_sort_and_sweep_pltp": [p, i32, i32, p, p, p, p, p, d, p],
        "swcu_encounter_check_all_sort_and_sweep_plplm": [p, i32, i32, p, p, p, p, p, p, d, p],
        "swcu_encounter_check_all_plplm": [p, i32, i32, p, p, p, p, p, p, d, p],
        "swcu_encounter_fetch": [p, i64, p, p, p],
        "swcu_encounter_stats": [p, p, p],
        "swcu_body_sync": [p, i32, i32, i32, p, p, p, p, p, p, p, u64],
        "swcu_body_put": [p, i32, p, p, p, p],
        "swcu_body_get": [p, i32, p, p, p, p],
        "swcu_body_put_range": [p, i32, i32, i32, p, p],
        "swcu_body_get_range": [p, i32, i32, i32, p, p, p],
        "swcu_body_put_range_async": [p, i32, i32, i32, p, p],
        "swcu_body_get_range_async": [p, i32, i32, i32, p, p, p],
        "swcu_io_wait": [p],
        "swcu_body_count": [p, i32, p, p, p],
        "swcu_body_zero_accel": [p, i32],
        "swcu_pl_accel_int": [p, i32, i32],
        "swcu_tp_accel_int": [p],
        "swcu_pl_set_renc": [p, i32],
        "swcu_body_drift": [p, i32, d, i32, d, p],
        "swcu_body_kick_velocity": [p, i32, d],
        "swcu_whm_tp_step": [p, d, p, p],
        "swcu_whm_step_pl": [p, d, d, i32, i32, i32, p],
        "swcu_whm_tp_first_accel": [p],
        "swcu_whm_get_jacobi": [p, p, p],
        "swcu_pl_vh2vb": [p, d, p],
        "swcu_pl_vb2vh": [p, d, p],
        "swcu_pl_lindrift": [p, d, d, i32, p],
        "swcu_tp_lindrift": [p, d, i32],
        "swcu_cb_set_pt": [p, p, p],
        "swcu_cb_get_pt": [p, p, p],
        "swcu_tp_vh2vb": [p, i32],
        "swcu_tp_vb2vh": [p, i32],
        "swcu_body_kick_vb": [p, i32, d, i32],
        "swcu_body_drift_vb": [p, i32, d, d, p],
        "swcu_body_put_vb": [p, i32, p],
        "swcu_body_set_active": [p, i32, p],
        "swcu_body_get_vb": [p, i32, p, p, p],
        "swcu_helio_step_pl": [p, d, d, i32, i32, i32, p],
        "swcu_helio_step_tp": [p, d, d, i32, p],
        "swcu_encounter_check_all_triangular_plpl": [p, i32, p, p, p, d, p],
        "swcu_encounter_check_all_triangular_pltp": [p, i32, i32, p, p, p, p, p, d, p],
        "swcu_encounter_check_all_triangular_plplm": [p, i32, i32, p, p, p, p, p, p, d, p],
        "swcu_discard_pl_tp": [p, i32, i32, p, p, p, p, p, p, d, p, p],
        "swcu_symba_encounter_check_list": [p, i64, p, p, p, i32, p, p, p, p, i32, p, p, p, p, d, p, p, p],
        "swcu_symba_kick_list_plpl": [p, i64, p, p, p, i32, p, p, p, p, d, i32, i32, p, p],
        "swcu_symba_kick_list_pltp": [p, i64, p, p, p, i32, i32, p, p, p, p, p, p, d, i32, i32, p, p],
        "swcu_collision_check_list": [p, i64, p, p, p, p, i32, p, p, p, p, i32, p, p, d, p, p, p],
        "swcu_util_get_potential_energy": [p, i32, p, d, p, p, p, p],
        "swcu_util_get_energy_and_momentum": [p, i32, p, d, d, p, p, p, p, p, p, p, i32, p],
        "swcu_pl_encounter_check": [p, d, p],
        "swcu_tp_encounter_check": [p, d, p],
        "swcu_comm_unique_id": [p, p],
        "swcu_comm_init": [p, i32, i32, p],
        "swcu_comm_finalize": [p],
        "swcu_pl_set_slice": [p, i32, i32],
        "swcu_pl_allgather": [p, i32],
        "swcu_partition": [i32, i32, i32, p, p],
        "swcu_p2p_export": [p, p],
        "swcu_p2p_import": [p, i32, i32, p],
        "swcu_p2p_close": [p],
        "swcu_pl_kick_drift_p2p": [p, i32, d, p],
        "swcu_timer_start": [p],
        "swcu_timer_stop": [p, p],
        "swcu_timer_lap_begin": [p],
        "swcu_timer_lap_end": [p],
        "swcu_timer_laps": [p, p, p, p, i32],
        "swcu_probe_fp64_peak": [p, p],
        "swcu_probe_hbm_copy": [p, i64, p],
        "swcu_flush_l2": [p],
        "swcu_last_kernel_ms": [p, i32, p],
        "swcu_enable_kernel_timing": [p, i32],
        "swcu_kernel_ms_accumulated": [p, i32, p, p],
        "swcu_flat_redo_count": [p, p],
        "swcu_encounter_direct_count": [p, p, p],
        "swcu_encounter_bucket_fallbacks": [p, p],
        "swcu_step_graph_replays": [p, p],
        "swcu_tp_discard_pl": [p, d, p, p],
        "swcu_pl_encounter_check_triangular": [p, d, p],
        "swcu_tp_encounter_check_triangular": [p, d, p],
        "swcu_pl_symba_kick_list": [p, i64, p, p, p, p, d, i32, i32, p],
        "swcu_tp_symba_kick_list": [p, i64, p, p, p, p, p, d, i32, i32, p],
        "swcu_body_symba_encounter_check_list": [p, i32, i64, p, p, p, d, p, p, p],
        "swcu_body_collision_check_list": [p, i32, i64, p, p, p, p, d, p, p, p],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.swcu_last_error.argtypes = [p]
    L.swcu_last_error.restype = C.c_char_p
    L.swcu_version.argtypes = []
    L.swcu_version.restype = C.c_int
    L.swcu_launch_count.argtypes = [p]
    L.swcu_launch_count.restype = i64
    _lib = L
    return L
