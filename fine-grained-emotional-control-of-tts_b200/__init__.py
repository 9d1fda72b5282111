"""fs2_b200: B200-native FastSpeech2 hot path (drop-in for emo_rank_tts/fastspeech2 model.py + loss.py)."""
from . import _lib  # noqa: F401
