"""Quick-start wrappers -- python_package/madflow/utilities.py (`one_matrix_integration` :56-90,
`generate_initial_states` :25-39)."""
from .integrand import FusedIntegrand
from .vegas import VegasFlow


def generate_initial_states(matrices):
    """utilities.py:25-39."""
    initial_flavours = []
    for matrix in matrices:
        initials = matrix.initial_states
        flavs_1, flavs_2 = zip(*initials)
        if matrix.mirror_initial_states:
            m2, m1 = zip(*initials)
            flavs_1 += m1
            flavs_2 += m2
        initial_flavours.append((flavs_1, flavs_2))
    return initial_flavours


def one_matrix_integration(matrix, model, sqrts=7e3, n_events=int(1e5), n_iter=5, q=91.46, pdf=None,
                           flavours=None, out_masses=None, alpha_s=0.118, seed=4, fused=True):
    """utilities.py:56-90: integrate one matrix element with ramboflow (COM-frame momenta, no cuts)
    and VEGAS.  `pdf` must be None: PDF interpolation needs an LHAPDF grid (pdfflow), which is outside
    this build (DESIGN.md "Out of scope")."""
    if pdf is not None:
        raise NotImplementedError("PDF luminosities need pdfflow + an LHAPDF grid; run with pdf=None")
    integrand = FusedIntegrand(matrix, model, sqrts=sqrts, masses=out_masses, lab_frame=False, alpha_s=alpha_s)
    vegas = VegasFlow(integrand.n_dim, n_events, seed=seed)
    vegas.compile(integrand if fused else integrand.python_integrand())
    return vegas.run_integration(n_iter)
