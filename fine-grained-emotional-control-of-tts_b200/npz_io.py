"""On-disk utterance format either side of the step: the `.npz` files written by
`/root/reference/emo_rank_tts/rank_model/preprocess.py:134-151` and read back by `FastSpeech2Dataset.__getitem__`
(`/root/reference/emo_rank_tts/fastspeech2/dataset.py:29-58`).

    metadata : phones (list of str), emotion, speaker, audio_id, audio_path, transcript, textgrid_path
    inputs   : mel (n_mels, T) f32, pitch (T,), energy (T,), durations (n_phones,) int      with sum(durations) == T

`read_npz_utterance` returns the sample dict of dataset.py:47-57, i.e. exactly what `DeviceCollate` (collate.py) and the
reference's `TextMelCollateWithAlignment` consume; `NpzUtterances` is the Dataset over a `fs2_{mode}.txt` path list
(dataset.py:20-27).  Host-side only: no GPU work happens here.

The phoneme -> id table is `['@'] + speechbrain.utils.text_to_sequence.valid_symbols + ['sil', 'spn', 'sp', '']`
(`fastspeech2/util.py:11-12, 30-32`).  speechbrain is absent from this image, so `DEFAULT_VALID_TOKENS` restates its
`valid_symbols` as the CMUdict ARPAbet list (84 symbols) -- UNPINNED; pass `valid_tokens=` to use the list of the
installation that wrote the files."""
from __future__ import annotations

import os

import numpy as np
import torch

_ARPABET = ("AA AA0 AA1 AA2 AE AE0 AE1 AE2 AH AH0 AH1 AH2 AO AO0 AO1 AO2 AW AW0 AW1 AW2 AY AY0 AY1 AY2 B CH D DH EH EH0 EH1 "
            "EH2 ER ER0 ER1 ER2 EY EY0 EY1 EY2 F G HH IH IH0 IH1 IH2 IY IY0 IY1 IY2 JH K L M N NG OW OW0 OW1 OW2 OY OY0 OY1 "
            "OY2 P R S SH T TH UH UH0 UH1 UH2 UW UW0 UW1 UW2 V W Y Z ZH").split()
SIL_PHONES = ["sil", "spn", "sp", ""]                       # util.py:11
DEFAULT_VALID_TOKENS = ["@"] + _ARPABET + SIL_PHONES        # util.py:12 (see the module docstring)

NPZ_KEYS = ("phones", "emotion", "speaker", "audio_id", "audio_path", "transcript", "textgrid_path",
            "mel", "pitch", "energy", "durations")            # preprocess.py:134-151


def phoneme2sequence(phones, valid_tokens=None):
    """util.py:30-32: index of every phone in the token table (ValueError for an unknown phone, as list.index)."""
    table = DEFAULT_VALID_TOKENS if valid_tokens is None else list(valid_tokens)
    lut = {}
    for i, tok in enumerate(table):
        lut.setdefault(tok, i)                 # list.index returns the FIRST occurrence
    try:
        return [lut[p] for p in phones]
    except KeyError as e:
        raise ValueError(f"{e.args[0]!r} is not in list") from None


def write_npz_utterance(path, *, phones, emotion, speaker, audio_id, audio_path, transcript, textgrid_path, mel, pitch,
                        energy, durations):
    """preprocess.py:134-151 (np.savez with the same keys; `phones` becomes an object/str array as numpy makes it)."""
    mel, pitch, energy = np.asarray(mel), np.asarray(pitch), np.asarray(energy)
    durations = np.asarray(durations)
    assert mel.shape[1] == len(pitch) == len(energy), "preprocess.py:133"
    np.savez(path, phones=phones, emotion=emotion, speaker=speaker, audio_id=audio_id, audio_path=audio_path,
             transcript=transcript, textgrid_path=textgrid_path, mel=mel, pitch=pitch, energy=energy, durations=durations)


def read_npz_utterance(path, speakers, emotions, noise_symbol="", valid_tokens=None):
    """dataset.py:29-58: one sample dict (mel (n_mels, T) f32, pitch / energy (T,) f32, duration / phoneme i64, speaker and
    emotion as 0-d index tensors, text with the noise symbol stripped, audio_path)."""
    data = np.load(path, allow_pickle=True)
    missing = [k for k in NPZ_KEYS if k not in data.files]
    if missing:
        raise KeyError(f"{path}: not a preprocess.py utterance file, missing {missing}")
    phoneme = data["phones"].tolist()
    speaker = data["speaker"].item()
    emotion = data["emotion"].item()
    text = data["transcript"].item().replace(noise_symbol.strip(), "").strip() if noise_symbol.strip() else \
        data["transcript"].item().strip()
    return {
        "mel": torch.FloatTensor(data["mel"]),
        "pitch": torch.FloatTensor(data["pitch"]),
        "energy": torch.FloatTensor(data["energy"]),
        "duration": torch.LongTensor(data["durations"]),
        "phoneme": torch.LongTensor(phoneme2sequence(phoneme, valid_tokens)),
        "speaker": torch.tensor(speakers.index(speaker), dtype=torch.long),
        "emotion": torch.tensor(emotions.index(emotion), dtype=torch.long),
        "text": text,
        "audio_path": data["audio_path"].item(),
    }


class NpzUtterances(torch.utils.data.Dataset):
    """dataset.py:12-27: the utterances listed in `<preprocessed_path>/fs2_<mode>.txt`, one .npz path per line."""

    def __init__(self, preprocessed_path, noise_symbol, speakers, emotions, mode="train", valid_tokens=None):
        self.noise_symbol, self.speakers, self.emotions, self.valid_tokens = noise_symbol, speakers, emotions, valid_tokens
        with open(os.path.join(preprocessed_path, f"fs2_{mode}.txt")) as f:
            self.data_paths = [line.strip() for line in f.readlines()]

    def __len__(self):
        return len(self.data_paths)

    def __getitem__(self, idx):
        return read_npz_utterance(self.data_paths[idx], self.speakers, self.emotions, self.noise_symbol, self.valid_tokens)
