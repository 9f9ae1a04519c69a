from . import op, stylesdf_model  # noqa: F401
