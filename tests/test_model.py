"""Host logic that mirrors the R helpers of the reference CLI (src/single_group/src/r/model_functions.R)."""
import numpy as np

from hygeia_b200 import model


def test_beta_parameters_known_answers():
    # SURVEY.md appendix D-3
    a, b = model.beta_parameters(model.DEFAULT_MU, model.DEFAULT_SIGMA)
    assert np.allclose(a, [17.1, 0.9, 12, 3, 12, 1.00000035951991], rtol=1e-12)
    assert np.allclose(b, [0.9, 17.1, 3, 12, 12, 1.00000035951991], rtol=1e-12)


def test_vartheta_layout():
    v, d = model.get_known_parameters(u=3)
    assert d == 36 and len(v) == 2 + 6 + 6 + 1 + 6
    assert v[0] == 3 and v[1] == 6 and v[14] == 1.0 and np.all(v[15:] == 2.0)


def test_theta_roundtrip_transposes_p():
    # model_functions.R:65-78 extracts p column-major while C++ reads blocks as rows
    rng = np.random.default_rng(0)
    p = rng.random((6, 6)); np.fill_diagonal(p, 0.0); p /= p.sum(1, keepdims=True)
    om = np.array(model.DEFAULT_OMEGA)
    th = model.convert_model_parameters_to_theta(p, om)
    p2, om2 = model.convert_theta_to_model_parameters(th)
    assert np.allclose(om2, om)
    pt = p.T.copy(); pt /= pt.sum(1, keepdims=True)
    assert np.allclose(p2, pt)
    # a symmetric P survives the round trip unchanged
    th = model.default_theta()
    p3, _ = model.convert_theta_to_model_parameters(th)
    assert np.allclose(p3, model.default_p())
