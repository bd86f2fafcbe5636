"""ctypes binding of libstefan_b200.so (the C ABI declared in include/stefan_b200.h).

The library is built in-tree by `stefanproblem_b200.build` (nvcc, sm_100a).  There
is no fallback of any kind: if the shared library is missing or a CUDA device is
not available the compute entry points raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# STEFAN_B200_LIB: another build of the same library (kernel tuning variants)
LIB_PATH = os.environ.get("STEFAN_B200_LIB") or os.path.join(_HERE, "lib", "libstefan_b200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int8_p = ctypes.POINTER(ctypes.c_int8)
c_int64_p = ctypes.POINTER(ctypes.c_int64)


class StefanError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libstefan_b200 status {status}: {message}")
        self.status = status


class IpdgParams(ctypes.Structure):
    """`stefan_ipdg_params` (include/stefan_b200.h)."""
    _fields_ = [
        ("nx", ctypes.c_int32), ("ny", ctypes.c_int32),
        ("hx", ctypes.c_double), ("hy", ctypes.c_double),
        ("order", ctypes.c_int32), ("numqp", ctypes.c_int32),
        ("k1", ctypes.c_double), ("k2", ctypes.c_double),
        ("penalty", ctypes.c_double),
        ("lam", ctypes.c_double), ("Tm", ctypes.c_double),
        ("variant", ctypes.c_int32), ("terms", ctypes.c_int32),
    ]


class CgParams(ctypes.Structure):
    """`stefan_cg_params`."""
    _fields_ = [("max_iterations", ctypes.c_int32), ("rel_tolerance", ctypes.c_double),
                ("preconditioner", ctypes.c_int32), ("check_every", ctypes.c_int32)]


class CgResult(ctypes.Structure):
    """`stefan_cg_result`."""
    _fields_ = [("iterations", ctypes.c_int32), ("residual_norm", ctypes.c_double), ("rhs_norm", ctypes.c_double),
                ("converged", ctypes.c_int32)]

# callbacks of stefan_ipdg_cg_solve_slab (include/stefan_b200.h: stefan_halo_fn, stefan_allreduce_fn)
HALO_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p)


class PeerSync(ctypes.Structure):
    """`stefan_peer_sync`."""
    _fields_ = [("my_ready", ctypes.c_void_p), ("my_done", ctypes.c_void_p), ("lo_ready", ctypes.c_void_p),
                ("hi_ready", ctypes.c_void_p), ("epoch", ctypes.c_uint64), ("timeout_ms", ctypes.c_int32),
                ("status", ctypes.c_void_p), ("epoch_dev", ctypes.c_void_p),
                ("post_ready", ctypes.c_void_p * 2), ("post_done", ctypes.c_void_p * 2), ("chained", ctypes.c_int32)]


class BlocksDims(ctypes.Structure):
    """`stefan_blocks_dims`."""
    _fields_ = [("ncells", ctypes.c_int32), ("order", ctypes.c_int32), ("nqv", ctypes.c_int32),
                ("nslots", ctypes.c_int32), ("nqf", ctypes.c_int32)]


class BlocksData(ctypes.Structure):
    """`stefan_blocks_data` (device pointers)."""
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "vol_pts", "vol_wts", "cell_jac", "cell_k", "slot_nbr", "face_pts", "face_nbr_pts", "face_wts",
        "face_normals", "slot_pen", "slot_lambda")]


class LdgParams(ctypes.Structure):
    """`stefan_ldg_params` (include/stefan_b200.h)."""
    _fields_ = [
        ("nx", ctypes.c_int32), ("ny", ctypes.c_int32),
        ("hx", ctypes.c_double), ("hy", ctypes.c_double),
        ("order", ctypes.c_int32), ("numqp", ctypes.c_int32),
        ("k1", ctypes.c_double), ("k2", ctypes.c_double),
        ("interiorpenalty", ctypes.c_double), ("negboundarypenalty", ctypes.c_double),
        ("posboundarypenalty", ctypes.c_double),
        ("V0x", ctypes.c_double), ("V0y", ctypes.c_double),
        ("interfacepenalty", ctypes.c_double), ("aligned_interface", ctypes.c_int32),
    ]


class Dg1dParams(ctypes.Structure):
    """`stefan_dg1d_params` (include/stefan_b200.h)."""
    _fields_ = [
        ("nelmts1", ctypes.c_int32), ("nelmts2", ctypes.c_int32),
        ("xL", ctypes.c_double), ("xR", ctypes.c_double), ("interfacepoint", ctypes.c_double),
        ("order", ctypes.c_int32), ("numqp", ctypes.c_int32),
        ("k1", ctypes.c_double), ("k2", ctypes.c_double), ("penalty", ctypes.c_double),
        ("lam", ctypes.c_double), ("Tm", ctypes.c_double),
        ("terms", ctypes.c_int32),
    ]


_SIGNATURES = {
    "stefan_last_error": (ctypes.c_char_p, []),
    "stefan_version": (ctypes.c_int, []),
    "stefan_reference_tables": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32] + [c_double_p] * 7),
    "stefan_ipdg_create": (ctypes.c_int, [ctypes.POINTER(IpdgParams), c_int8_p, c_int8_p, c_int8_p,
                                          ctypes.POINTER(ctypes.c_void_p)]),
    "stefan_ipdg_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "stefan_ipdg_ndofs": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_ldg_create": (ctypes.c_int, [ctypes.POINTER(LdgParams), c_int8_p, c_int8_p, c_int8_p,
                                         ctypes.POINTER(ctypes.c_void_p)]),
    "stefan_ldg_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "stefan_ldg_ndofs": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_ldg_coo_size": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_ldg_assemble_coo": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_int64, ctypes.c_int32, c_int64_p, ctypes.c_void_p]),
    "stefan_dg1d_coo_size": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_dg1d_assemble_coo": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_int64, ctypes.c_int32, c_int64_p, ctypes.c_void_p]),
    "stefan_ldg_rhs": (ctypes.c_int, [ctypes.c_void_p] * 5),
    "stefan_ldg_apply": (ctypes.c_int, [ctypes.c_void_p] * 6),
    "stefan_ldg_apply_algo": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int32, ctypes.c_void_p]),
    "stefan_ldg_solve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.POINTER(CgParams), ctypes.POINTER(CgResult), ctypes.c_void_p]),
    "stefan_ipdg_cg_solve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.POINTER(CgParams), ctypes.POINTER(CgResult), ctypes.c_void_p]),
    "stefan_ipdg_cg_solve_slab": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.POINTER(CgParams), ctypes.POINTER(CgResult), ctypes.c_void_p,
                                                 ctypes.c_void_p, HALO_FN, ALLREDUCE_FN, ctypes.c_void_p, ctypes.c_void_p]),
    "stefan_ipdg_l2_error": (ctypes.c_int, [ctypes.c_void_p] * 5),
    "stefan_ipdg_apply_peer": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.POINTER(PeerSync), ctypes.c_void_p]),
    "stefan_ipdg_step": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.POINTER(PeerSync), ctypes.c_void_p]),
    "stefan_ipdg_interface_sums_peer": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.POINTER(PeerSync), ctypes.c_void_p,
                                                       ctypes.c_void_p]),
    "stefan_blocks_create": (ctypes.c_int, [ctypes.POINTER(BlocksDims), ctypes.POINTER(ctypes.c_void_p)]),
    "stefan_blocks_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "stefan_blocks_ndofs": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_blocks_set_variant": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "stefan_blocks_assemble": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(BlocksData), ctypes.c_void_p]),
    "stefan_blocks_apply": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "stefan_blocks_coo_size": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_blocks_assemble_coo": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_int64, ctypes.c_int32, c_int64_p, ctypes.c_void_p]),
    "stefan_blocks_rhs": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(BlocksData), ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "stefan_blocks_l2_error": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(BlocksData)] + [ctypes.c_void_p] * 4),
    "stefan_blocks_cg_solve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.POINTER(CgParams), ctypes.POINTER(CgResult), ctypes.c_void_p]),
    "stefan_ldgblocks_create": (ctypes.c_int, [ctypes.POINTER(BlocksDims), ctypes.POINTER(ctypes.c_void_p)]),
    "stefan_ldgblocks_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "stefan_ldgblocks_ndofs": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_ldgblocks_assemble": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(BlocksData), ctypes.c_double,
                                                 ctypes.c_double, ctypes.c_void_p]),
    "stefan_ldgblocks_apply": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "stefan_ldgblocks_rhs": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(BlocksData)] + [ctypes.c_void_p] * 4),
    "stefan_ldgblocks_coo_size": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_ldgblocks_assemble_coo": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                     ctypes.c_int64, ctypes.c_int32, c_int64_p, ctypes.c_void_p]),
    "stefan_ldgblocks_solve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.POINTER(CgParams), ctypes.POINTER(CgResult), ctypes.c_void_p]),
    "stefan_blocks_interface_l2_error": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(BlocksData)] + [ctypes.c_void_p] * 5),
    "stefan_interpolate_at_reference_points": (ctypes.c_int, [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                                              ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                                              ctypes.c_void_p]),
    "stefan_gradient_at_reference_points": (ctypes.c_int, [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                                           ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                                           ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "stefan_front_velocity": (ctypes.c_int, [ctypes.c_int64] + [ctypes.c_double] * 4 + [ctypes.c_void_p] * 10),
    "stefan_interpolate_field_at_reference_points": (ctypes.c_int, [ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                                                    ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                                                    ctypes.c_void_p]),
    "stefan_ldg_interface_flux": (ctypes.c_int, [ctypes.c_int64] + [ctypes.c_double] * 4 + [ctypes.c_void_p] * 7),
    "stefan_mesh_l2_error": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_double,
                                            ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "stefan_peer_alloc": (ctypes.c_int, [ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    "stefan_peer_free": (ctypes.c_int, [ctypes.c_void_p]),
    "stefan_peer_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "stefan_peer_close": (ctypes.c_int, [ctypes.c_void_p]),
    "stefan_peer_signal": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]),
    "stefan_peer_wait": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int32, ctypes.c_void_p,
                                        ctypes.c_void_p]),
    "stefan_dg1d_create": (ctypes.c_int, [ctypes.POINTER(Dg1dParams), ctypes.POINTER(ctypes.c_void_p)]),
    "stefan_dg1d_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "stefan_dg1d_ndofs": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_dg1d_apply": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "stefan_dg1d_rhs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "stefan_fd_create": (ctypes.c_int, [ctypes.c_int64, ctypes.c_double, c_int8_p, ctypes.c_int32, c_int8_p,
                                        c_int8_p, ctypes.POINTER(ctypes.c_void_p)]),
    "stefan_fd_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "stefan_fd_add_interface_continuity": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                                          ctypes.c_int64]),
    "stefan_fd_apply": (ctypes.c_int, [ctypes.c_void_p] * 6),
    "stefan_fd_step": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_double] * 3 + [ctypes.c_void_p]),
    "stefan_fd_nnz": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_fd_assemble_coo": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int64, c_int64_p, ctypes.c_void_p]),
    "stefan_ipdg_apply_columns": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int32] * 3 + [ctypes.c_void_p]),
    "stefan_ipdg_interface_sums": (ctypes.c_int, [ctypes.c_void_p] * 6),
    "stefan_ipdg_euler_update": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_double, ctypes.c_void_p]),
    "stefan_ipdg_rhs": (ctypes.c_int, [ctypes.c_void_p] * 5),
    "stefan_ipdg_coo_size": (ctypes.c_int64, [ctypes.c_void_p]),
    "stefan_ipdg_assemble_coo": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int32,
                                                                          c_int64_p, ctypes.c_void_p]),
    "stefan_ipdg_apply_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int32]),
    "stefan_ipdg_apply": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                         ctypes.c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StefanError(-1, f"{LIB_PATH} not found: run `python -m stefanproblem_b200.build` "
                                  "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise StefanError(status, lib().stefan_last_error().decode())
