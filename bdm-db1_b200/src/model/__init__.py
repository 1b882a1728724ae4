from .transformer_xl import TransformerXL  # noqa: F401
