"""`InteriorPenalty` module of the reference (StefanDG/InteriorPenalty/interior_penalty.jl),
same function names and argument order, over the CUDA path.

Each `assemble_*` call records its pass in the `SystemMatrix` description; the
operator itself is applied matrix-free by libstefan_b200.so after
`CutCellDG.sparse_operator` (see cutcelldg.py).
"""
from . import cutcelldg as _cc

VOLUME, INTERELEMENT_FLUX, INTERELEMENT_PENALTY = 1, 2, 4
INTERFACE_FLUX, INTERFACE_PENALTY, BOUNDARY_FLUX, BOUNDARY_PENALTY, ROBIN = 8, 16, 32, 64, 128
REFERENCE_SYSTEM = VOLUME | INTERELEMENT_FLUX | INTERELEMENT_PENALTY | BOUNDARY_FLUX | BOUNDARY_PENALTY


def _common(basis, quads):
    return dict(order=basis.order, numqp=quads.numqp)


def assemble_gradient_operator(sysmatrix, basis, cellquads, k1, k2, mesh):
    """IP_gradient_operator.jl:29-57"""
    sysmatrix._merge("ip", terms=VOLUME, k1=float(k1), k2=float(k2), **_common(basis, cellquads))


def assemble_interelement_flux_operator(sysmatrix, basis, facequads, k1, k2, mesh):
    """IP_flux_operator.jl:202-256"""
    sysmatrix._merge("ip", terms=INTERELEMENT_FLUX, k1=float(k1), k2=float(k2),
                     **_common(basis, facequads))


def assemble_interelement_penalty_operator(sysmatrix, basis, facequads, penalty, mesh):
    """IP_penalty_operator.jl:134-181"""
    sysmatrix._merge("ip", terms=INTERELEMENT_PENALTY, penalty=float(penalty),
                     **_common(basis, facequads))


def assemble_interface_flux_operator(sysmatrix, basis, interfacequads, k1, k2, mesh):
    """IP_flux_operator.jl:296-324: cut cells only -- none on an uncut mesh."""


def assemble_interface_penalty_operator(sysmatrix, basis, interfacequads, penalty, mesh):
    """IP_penalty_operator.jl:213-237: cut cells only -- none on an uncut mesh."""


def assemble_boundary_flux_operator(sysmatrix, basis, facequads, k1, k2, mesh, variant="current"):
    """IP_flux_operator.jl:393-444"""
    sysmatrix._merge("ip", terms=BOUNDARY_FLUX, k1=float(k1), k2=float(k2), variant=variant,
                     **_common(basis, facequads))


def assemble_boundary_penalty_operator(sysmatrix, basis, facequads, penalty, mesh):
    """IP_penalty_operator.jl:286-329"""
    sysmatrix._merge("ip", terms=BOUNDARY_PENALTY, penalty=float(penalty),
                     **_common(basis, facequads))


def assemble_interior_penalty_linear_system(sysmatrix, solverbasis, cellquads, facequads,
                                            interfacequads, k1, k2, penalty, mesh,
                                            variant="current"):
    """IP_assembly.jl:1-80, same order of passes."""
    assemble_gradient_operator(sysmatrix, solverbasis, cellquads, k1, k2, mesh)
    assemble_interelement_flux_operator(sysmatrix, solverbasis, facequads, k1, k2, mesh)
    assemble_interelement_penalty_operator(sysmatrix, solverbasis, facequads, penalty, mesh)
    assemble_interface_flux_operator(sysmatrix, solverbasis, interfacequads, k1, k2, mesh)
    assemble_interface_penalty_operator(sysmatrix, solverbasis, interfacequads, penalty, mesh)
    assemble_boundary_flux_operator(sysmatrix, solverbasis, facequads, k1, k2, mesh, variant)
    assemble_boundary_penalty_operator(sysmatrix, solverbasis, facequads, penalty, mesh)


# ---- mesh-aligned interface: the reference's face-level interface functions applied
# ---- to the faces between uncut cells of unlike phase
def assemble_aligned_interface_flux_operator(sysmatrix, basis, facequads, k1, k2, mesh):
    """assemble_face_flux_operator! (IP_flux_operator.jl:35-118) on unlike-phase faces."""
    sysmatrix._merge("ip", terms=INTERFACE_FLUX, k1=float(k1), k2=float(k2),
                     **_common(basis, facequads))


def assemble_aligned_interface_penalty_operator(sysmatrix, basis, facequads, penalty, mesh):
    """assemble_face_penalty_operator! (IP_penalty_operator.jl:21-65) on unlike-phase faces."""
    sysmatrix._merge("ip", terms=INTERFACE_PENALTY, penalty=float(penalty),
                     **_common(basis, facequads))


def assemble_robin_operator(sysmatrix, basis, interfacequads_or_facequads, lam, mesh):
    """IP_robin_interface_operator.jl:76-99 (face-level form on unlike-phase faces)."""
    sysmatrix._merge("ip", terms=ROBIN, lam=float(lam))


def assemble_boundary_source(sysrhs, rhsfunc, basis, facequads, k1, k2, penalty, mesh, variant="current"):
    """IP_source_term.jl:201-258 (`rhsfunc(x)` is called with x = [xs; ys], 2 x N)."""
    sysrhs._merge("ip", g=rhsfunc, k1=float(k1), k2=float(k2), penalty=float(penalty), variant=variant,
                  **_common(basis, facequads))


def assemble_source(sysrhs, rhsfunc, basis, cellquads, mesh):
    """IP_source_term.jl:19-49"""
    assemble_two_phase_source(sysrhs, rhsfunc, rhsfunc, basis, cellquads, mesh)


def assemble_two_phase_source(sysrhs, rhsfunc1, rhsfunc2, basis, cellquads, mesh):
    """IP_source_term.jl:54-91"""
    sysrhs._merge("ip", f1=rhsfunc1, f2=rhsfunc2, **_common(basis, cellquads))


def assemble_robin_source(sysrhs, basis, interfacequads_or_facequads, lam, Tm, mesh):
    """IP_robin_interface_source.jl:65-83 (face-level form on unlike-phase faces)."""
    sysrhs._merge("ip", lam=float(lam), Tm=float(Tm), order=basis.order,
                  numqp=interfacequads_or_facequads.numqp)


def assemble_interior_penalty_rhs(sysrhs, sourceterm, boundaryfunc, solverbasis, cellquads, facequads,
                                  k1, k2, penalty, mesh, variant="current"):
    """IP_assembly.jl:82-115"""
    assemble_boundary_source(sysrhs, boundaryfunc, solverbasis, facequads, k1, k2, penalty, mesh, variant)
    assemble_source(sysrhs, sourceterm, solverbasis, cellquads, mesh)


sparse_operator = _cc.sparse_operator
rhs_vector = _cc.rhs_vector
