"""ctypes bindings of include/uppasd_b200.h (argument types only -- no logic lives here)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ASD_LIB') or os.path.join(_HERE, 'libuppasd_b200.so')   # ASD_LIB: development builds

c_int_p = C.POINTER(C.c_int)
c_uint_p = C.POINTER(C.c_uint)
c_dbl_p = C.POINTER(C.c_double)
vp = C.c_void_p

CB_DO = C.CFUNCTYPE(None, C.POINTER(C.c_size_t), c_int_p)
CB_MEASURE = C.CFUNCTYPE(None, c_dbl_p, c_dbl_p, c_dbl_p, C.POINTER(C.c_size_t))
CB_FLUSH = C.CFUNCTYPE(None, C.POINTER(C.c_size_t))
CB_STATUS = C.CFUNCTYPE(None, c_dbl_p)

# every symbol include/uppasd_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    'fortrandata_setconstants_': (None, [vp] * 21),
    'fortrandata_setmatrices_': (None, [vp] * 24),
    'fortrandata_setinputdata_': (None, [vp] * 3),
    'fortrandata_setextras_': (None, [vp] * 8),
    'fortrandata_setlattice_': (None, [vp] * 7),
    'cudamdsim_initiateconstants_': (None, []),
    'cudamdsim_initiatematrices_': (None, []),
    'cudamdsim_measurementphase_': (None, []),
    'asd_legacy_async_samples': (C.c_long, []),
    'cudamdsim_initialphase_': (None, [vp] * 6),
    'cudamcsim_evolve_': (None, [vp] * 7),
    'relax_': (None, [vp] * 8),
    'get_emom_': (None, [vp] * 3),
    'put_emom_': (None, [vp] * 3),
    'get_beff_': (None, [vp] * 3),
    'get_energy_': (None, [vp]),
    'cmdsim_initiateconstants_': (None, []),
    'cmdsim_initiatefortran_': (None, []),
    'cmdsim_measurementphase_': (None, []),
    'asd_set_callbacks': (None, [CB_DO, CB_MEASURE, CB_FLUSH, CB_STATUS]),
    'asd_legacy_engine': (vp, []),
    'asd_last_error': (C.c_char_p, []),
    'asd_device_count': (C.c_int, []),
    'asd_create': (C.c_int, [C.POINTER(vp), C.c_int]),
    'asd_destroy': (None, [vp]),
    'asd_set_constants': (C.c_int, [vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    'asd_set_system': (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp]),
    'asd_set_exchange': (C.c_int, [vp, C.c_int, vp, vp, vp]),
    'asd_set_jtensor': (C.c_int, [vp, C.c_int, vp, vp, vp]),
    'asd_measure_sublattice': (C.c_int, [vp, C.c_int, vp]),
    'asd_set_triangulation': (C.c_int, [vp, C.c_int, vp]),
    'asd_skyrmion_number': (C.c_int, [vp, vp]),
    'asd_set_evolving_atoms': (C.c_int, [vp, C.c_int, vp]),
    'asd_set_dm': (C.c_int, [vp, C.c_int, vp, vp, vp]),
    'asd_set_bq': (C.c_int, [vp, C.c_int, vp, vp, vp]),
    'asd_set_lattice_hint': (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]),
    'asd_set_anisotropy': (C.c_int, [vp, vp, vp, vp, vp]),
    'asd_set_external_field': (C.c_int, [vp, vp]),
    'asd_set_torque': (C.c_int, [vp, vp]),
    'asd_set_time_field': (C.c_int, [vp, C.c_long, C.c_long, vp]),
    'asd_set_llg': (C.c_int, [vp, C.c_int, C.c_double, vp, vp, vp, C.c_double, C.c_int, C.c_ulonglong]),
    'asd_set_moments': (C.c_int, [vp, vp, vp, vp]),
    'asd_get_moments': (C.c_int, [vp, vp, vp, vp]),
    'asd_commit': (C.c_int, [vp]),
    'asd_effective_field': (C.c_int, [vp, vp, vp, vp, vp]),
    'asd_sd_steps': (C.c_int, [vp, C.c_long, C.c_long]),
    'asd_sd_run': (C.c_int, [vp, C.c_long, C.c_long, C.c_long, vp, C.POINTER(C.c_long)]),
    'asd_mc_sweeps': (C.c_int, [vp, C.c_char, C.c_long, C.c_long, C.c_double, C.c_double, vp]),
    'asd_set_mc_layout': (C.c_int, [vp, C.c_int]),
    'asd_mc_colouring': (C.c_int, [vp, c_int_p, c_int_p, c_int_p]),
    'asd_get_mc_colours': (C.c_int, [vp, vp]),
    'asd_get_mc_visit_order': (C.c_int, [vp, vp]),
    'asd_debug_mc_draws': (C.c_int, [vp, C.c_long, vp, vp]),
    'asd_measure': (C.c_int, [vp, vp, vp]),
    'asd_energy_terms': (C.c_int, [vp, vp]),
    'asd_get_atoms': (C.c_int, [vp, C.c_int, vp, vp]),
    'asd_time_sd_steps': (C.c_int, [vp, C.c_long, C.c_long, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    'asd_time_mc_sweeps': (C.c_int, [vp, C.c_char, C.c_long, C.c_double, C.POINTER(C.c_float)]),
    'asd_layout_info': (C.c_int, [vp, c_int_p]),
    'asd_launch_count': (C.c_long, [vp]),
    'asd_synchronize': (C.c_int, [vp]),
    'asd_build_lattice_table': (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, vp, vp, vp, vp]),
    'asd_get_table_dims': (C.c_int, [vp, C.c_int, c_int_p, c_int_p]),
    'asd_get_table': (C.c_int, [vp, C.c_int, vp, vp, vp]),
    'asd_init_moments_tilted': (C.c_int, [vp, C.c_double, C.c_int, vp]),
    'asd_set_ensemble_offset': (C.c_int, [vp, C.c_uint]),
    'asd_set_slab': (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
    'asd_slab_handle_bytes': (C.c_int, []),
    'asd_slab_export': (C.c_int, [vp, vp]),
    'asd_slab_connect_ipc': (C.c_int, [vp, vp, vp]),
    'asd_slab_connect_local': (C.c_int, [vp, vp, vp]),
    'asd_slab_status': (C.c_int, [vp, C.POINTER(C.c_ulonglong), c_int_p]),
}

_lib = None


def load():
    """Loads the in-tree library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('libuppasd_b200.so is not built: run `python -m uppasd_b200.build` '
                               '(or __graft_entry__.build()); there is no CPU fallback')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
