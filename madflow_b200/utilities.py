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
    """utilities.py:56-90: integrate one matrix element with ramboflow (COM-frame momenta, no cuts) and VEGAS.
    pdf (madflow_b200.pdf.PDF, pdfflow's mkPDF) + flavours: the luminosity of the reference's `_generate_luminosity`
    (utilities.py:42-53),  prod_h x f_flavours(x_h, q^2) / x_h  at the FIXED scale q, for every flavour of `flavours`
    (0 = gluon); the couplings stay those of the frozen model (alpha_s), as in the reference where `model_params` is
    evaluated once.  This is the reference's numeric regression harness (tests/test_integration.py: 103.4 pb for
    g g > t t~ with NNPDF31_nnlo_as_0118 -- the grid itself is not available offline)."""
    states = None
    if pdf is not None:
        if flavours is None:
            raise ValueError("one_matrix_integration: pdf needs the flavours of the two incoming partons, e.g. (0,)")
        flavours = [flavours] if isinstance(flavours, int) else list(flavours)
        states = [[int(f), int(f)] for f in flavours]
    integrand = FusedIntegrand(matrix, model, sqrts=sqrts, masses=out_masses, lab_frame=False, alpha_s=alpha_s, pdf=pdf,
                               fixed_scale=q if pdf is not None else None, initial_states=states,
                               mirror_initial_states=False if pdf is not None else None)
    vegas = VegasFlow(integrand.n_dim, n_events, seed=seed)
    vegas.compile(integrand if fused else integrand.python_integrand())
    return vegas.run_integration(n_iter)
