from .speecht5_criterion import SpeechT5Criterion, SpeechT5CriterionConfig  # noqa: F401
from .text_to_speech_loss import TexttoSpeechLoss  # noqa: F401
from .speech_to_text_loss import SpeechtoTextLoss  # noqa: F401
from .speech_pretrain_criterion import SpeechPretrainCriterion  # noqa: F401
from .text_pretrain_criterion import TextPretrainCriterion  # noqa: F401
