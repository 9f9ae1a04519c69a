from . import encoders, helper_modules, op, stylesdf_model  # noqa: F401
