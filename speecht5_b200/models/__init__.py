from .speecht5 import T5TransformerModel, make_args  # noqa: F401
