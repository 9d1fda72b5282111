"""fs2_b200: B200-native FastSpeech2 hot path (drop-in for emo_rank_tts/fastspeech2 model.py + loss.py)."""
from . import _lib  # noqa: F401
from .model import FastSpeech2  # noqa: F401
from .loss import Loss  # noqa: F401
from .optim import FusedAdamW  # noqa: F401
from .intensity import get_intensity_representation, intensity_prototypes, intensity_segment_mean  # noqa: F401
from .rank_model import IntensityExtractor  # noqa: F401
from .collate import DeviceCollate  # noqa: F401
from .npz_io import NpzUtterances, read_npz_utterance, write_npz_utterance  # noqa: F401

DEFAULT_MODEL_CONFIG = dict(
    enc_num_layers=6, enc_num_head=2, enc_d_model=384, enc_ffn_dim=1536, enc_k_dim=384, enc_v_dim=384,
    enc_dropout=0.1, dec_num_layers=6, dec_num_head=2, dec_d_model=384, dec_ffn_dim=1536, dec_k_dim=384,
    dec_v_dim=384, dec_dropout=0.1, normalize_before=False, ffn_type="1dcnn", ffn_cnn_kernel_size_list=[9, 1],
    n_char=95, n_mels=80, postnet_embedding_dim=512, postnet_kernel_size=5, postnet_n_convolutions=5,
    postnet_dropout=0.5, padding_idx=0, dur_pred_kernel_size=3, pitch_pred_kernel_size=3,
    energy_pred_kernel_size=3, variance_predictor_dropout=0.5,
)  # /root/reference/emo_rank_tts/fastspeech2/parameter.yaml:62-90

DEFAULT_RANK_MODEL_CONFIG = dict(n_mels=80, n_heads=2, n_emotions=5, n_encoder_layers=6, hidden_dim=384, kernel_size=9,
                                 dropout=0.1)  # /root/reference/emo_rank_tts/rank_model/parameter.yaml:52-58

DEFAULT_LOSS_CONFIG = dict(
    log_scale_durations=True, ssim_loss_weight=1.0, duration_loss_weight=1.0, pitch_loss_weight=1.0,
    energy_loss_weight=1.0, mel_loss_weight=1.0, postnet_mel_loss_weight=1.0, spn_loss_weight=0.0,
    spn_loss_max_epochs=1,
)  # parameter.yaml:96-106
