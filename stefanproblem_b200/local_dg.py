"""`LocalDG` module of the reference (StefanDG/LocalDG/local_DG.jl), same function
names and argument order, over the CUDA path (see cutcelldg.py for how `SystemMatrix`
and `sparse_operator` work here)."""
from . import cutcelldg as _cc

DIVERGENCE, MASS, GRADIENT, SCALAR_FLUX, VECTOR_FLUX, BOUNDARY_VECTOR_FLUX = 1, 2, 4, 8, 16, 32


def _common(basis, quads):
    return dict(order=basis.order, numqp=quads.numqp)


def assemble_divergence_operator(sysmatrix, basis, cellquads, mesh):
    """LDG_div_operator.jl:44-70"""
    sysmatrix._merge("ldg", terms=DIVERGENCE, **_common(basis, cellquads))


def assemble_mass_operator(sysmatrix, basis, cellquads, mesh):
    """LDG_mass_operator.jl:27-53"""
    sysmatrix._merge("ldg", terms=MASS, **_common(basis, cellquads))


def assemble_gradient_operator(sysmatrix, basis, cellquads, k1, k2, mesh):
    """LDG_gradient_operator.jl:40-68"""
    sysmatrix._merge("ldg", terms=GRADIENT, k1=float(k1), k2=float(k2), **_common(basis, cellquads))


def assemble_interelement_scalar_flux_operator(sysmatrix, basis, facequads, V0, mesh):
    """LDG_scalar_flux_operator.jl:187-237"""
    sysmatrix._merge("ldg", terms=SCALAR_FLUX, V0=(float(V0[0]), float(V0[1])), **_common(basis, facequads))


def assemble_interelement_vector_flux_operator(sysmatrix, basis, facequads, k1, k2, alpha, beta, mesh):
    """LDG_vector_flux_operator.jl:243-300; `beta` is what the reference passes here: V0."""
    sysmatrix._merge("ldg", terms=VECTOR_FLUX, k1=float(k1), k2=float(k2), interiorpenalty=float(alpha),
                     V0=(float(beta[0]), float(beta[1])), **_common(basis, facequads))


def assemble_interface_scalar_flux_operator(sysmatrix, basis, interfacequads, V0, mesh):
    """LDG_scalar_flux_operator.jl:285-309: cut cells only -- none on an uncut mesh."""


def assemble_interface_vector_flux_operator(sysmatrix, basis, interfacequads, k1, k2, alpha, beta, mesh):
    """LDG_vector_flux_operator.jl:357-386: cut cells only -- none on an uncut mesh."""


def assemble_boundary_vector_flux_operator(sysmatrix, basis, facequads, k1, k2, negboundarypenalty,
                                           posboundarypenalty, V0, mesh):
    """LDG_vector_flux_operator.jl:478-535"""
    sysmatrix._merge("ldg", terms=BOUNDARY_VECTOR_FLUX, k1=float(k1), k2=float(k2),
                     negboundarypenalty=float(negboundarypenalty),
                     posboundarypenalty=float(posboundarypenalty), V0=(float(V0[0]), float(V0[1])),
                     **_common(basis, facequads))


def assemble_LDG_linear_system(sysmatrix, solverbasis, cellquads, facequads, interfacequads, k1, k2,
                               interiorpenalty, interfacepenalty, negboundarypenalty, posboundarypenalty,
                               V0, mesh):
    """LDG_assembly.jl:1-91, same order of passes (V0 handed on as `beta`, :44-53)."""
    assemble_divergence_operator(sysmatrix, solverbasis, cellquads, mesh)
    assemble_mass_operator(sysmatrix, solverbasis, cellquads, mesh)
    assemble_gradient_operator(sysmatrix, solverbasis, cellquads, k1, k2, mesh)
    assemble_interelement_scalar_flux_operator(sysmatrix, solverbasis, facequads, V0, mesh)
    assemble_interelement_vector_flux_operator(sysmatrix, solverbasis, facequads, k1, k2,
                                               interiorpenalty, V0, mesh)
    assemble_interface_scalar_flux_operator(sysmatrix, solverbasis, interfacequads, V0, mesh)
    assemble_interface_vector_flux_operator(sysmatrix, solverbasis, interfacequads, k1, k2,
                                            interfacepenalty, V0, mesh)
    assemble_boundary_vector_flux_operator(sysmatrix, solverbasis, facequads, k1, k2,
                                           negboundarypenalty, posboundarypenalty, V0, mesh)


def assemble_LDG_rhs(sysrhs, sourceterm, boundaryfunc, solverbasis, cellquads, facequads,
                     negboundarypenalty, posboundarypenalty, V0, mesh):
    """LDG_assembly.jl:93-131 (`sourceterm(x)`, `boundaryfunc(x)` are called with x = [xs; ys])."""
    sysrhs._merge("ldg", f=sourceterm, g=boundaryfunc, order=solverbasis.order, numqp=cellquads.numqp,
                  negboundarypenalty=float(negboundarypenalty), posboundarypenalty=float(posboundarypenalty),
                  V0=(float(V0[0]), float(V0[1])))


# ---- mesh-aligned interface: the reference's face-level interface functions applied to the faces between uncut
# ---- cells of unlike phase; both must be called (the kernel applies them together)
def assemble_aligned_interface_scalar_flux_operator(sysmatrix, basis, facequads, V0, mesh):
    """assemble_face_scalar_flux_operator! (LDG_scalar_flux_operator.jl:25-107) on unlike-phase faces, both
    orientations (as :241-283 does for the two sides of a cut cell's interface)."""
    sysmatrix._merge("ldg", V0=(float(V0[0]), float(V0[1])), **_common(basis, facequads))
    sysmatrix.spec["aligned_scalar"] = True


def assemble_aligned_interface_vector_flux_operator(sysmatrix, basis, facequads, k1, k2, alpha, beta, mesh):
    """assemble_face_vector_flux_operator! (LDG_vector_flux_operator.jl:47-154) on unlike-phase faces,
    alpha = interfacepenalty (as :318-355)."""
    sysmatrix._merge("ldg", k1=float(k1), k2=float(k2), interfacepenalty=float(alpha),
                     V0=(float(beta[0]), float(beta[1])), **_common(basis, facequads))


sparse_operator = _cc.sparse_operator
rhs_vector = _cc.rhs_vector
