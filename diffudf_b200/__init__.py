"""diffudf_b200 — B200-native hot path of DUDF (SIREN jets, hyperbolic UDF / Eikonal training step,
grid / ray / point field queries) behind the reference's Python call surface.  See DESIGN.md."""
from .model import SIREN, SineLayer  # noqa: F401
from .diff_operators import gradient, hessian, jacobian, laplace, divergence  # noqa: F401
from .loss_functions import loss_s1, loss_s2, loss_siren  # noqa: F401
from .evaluate import evaluate  # noqa: F401
from .inverses import inverse  # noqa: F401
