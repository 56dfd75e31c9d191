"""oracle/shim_tf/tensorflow_probability -- TEST INFRASTRUCTURE (see ../tensorflow/__init__.py).

The slice of TFP 0.11 the reference's two-group modules touch, on NumPy/SciPy.  Formulas restated from the TFP 0.11 sources as
documented (negative_binomial.py, beta_binomial.py, categorical.py, distribution.py: log_survival_function = log1p(-cdf)):
    NegativeBinomial.log_prob(x) = total_count log_sigmoid(-logits) + x log_sigmoid(logits) - lbeta(1 + x, total_count) - log(total_count + x)
    NegativeBinomial.cdf(x)      = betainc(total_count, 1 + x, sigmoid(-logits))
    BetaBinomial.log_prob(x)     = lbeta(c1 + x, c0 + n - x) - lbeta(c1, c0) + log C(n, x)
Values are computed in fp64 and rounded to the dtype TensorFlow would hold (fp32 in the reference's model).
"""
from . import python  # noqa: F401
from .python import distributions, bijectors, math  # noqa: F401
from .python.util import SeedStream  # noqa: F401

__version__ = "0.11.0-shim"
