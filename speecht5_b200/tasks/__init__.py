from .speecht5 import SpeechT5Task  # noqa: F401
