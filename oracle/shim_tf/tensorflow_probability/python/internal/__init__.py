from . import prefer_static, reparameterization, samplers  # noqa: F401
