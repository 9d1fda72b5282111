"""Runs the REAL reference glue -- emo_rank_tts/fastspeech2/model.py (FastSpeech2.__init__ / forward, model.py:149-441)
and fastspeech2/loss.py (Loss, loss.py:6-186) -- from where it lies under /root/reference, with the speechbrain leaf
symbols it imports (model.py:13-27, loss.py:3) supplied by the oracle's restatements.

speechbrain itself is absent from this image, so the leaves stay a restatement; but everything the
reference's own files do -- module construction and naming (hence the state_dict layout), the mask quirk, the
conditioning cat/projection, the variance adaptor order, the loss slicing and weighting -- is then the reference's
code, executed unmodified, and the oracle's FastSpeech2 / Loss glue can be checked against it instead of against a
reading of it.  Test infrastructure only (imports oracle/); needs /root/reference, i.e. the build container.
"""
import importlib.util
import os
import sys
import types

import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fs2_oracle as O  # noqa: E402

REF_DIR = "/root/reference/emo_rank_tts/fastspeech2"


def available():
    return os.path.isfile(os.path.join(REF_DIR, "model.py")) and os.path.isfile(os.path.join(REF_DIR, "loss.py"))


class _Conv1d(O.SBConv1d):
    """speechbrain.nnet.CNN.Conv1d call signature (model.py:226-240)."""

    def __init__(self, in_channels, out_channels, kernel_size, padding="same", skip_transpose=False):
        assert padding == "same"
        super().__init__(in_channels, out_channels, kernel_size, skip_transpose=skip_transpose)


class _TransformerEncoder(O.TransformerEncoder):
    """speechbrain TransformerEncoder call signature (model.py:241-267)."""

    def __init__(self, num_layers, nhead, d_ffn, d_model=None, kdim=None, vdim=None, dropout=0.0, activation=nn.ReLU,
                 normalize_before=False, ffn_type="regularFFN", ffn_cnn_kernel_size_list=(3, 3)):
        assert activation is nn.ReLU and ffn_type == "1dcnn"
        super().__init__(num_layers, nhead, d_ffn, d_model, kdim, vdim, dropout, normalize_before,
                         list(ffn_cnn_kernel_size_list))


def _stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    cnn = mod("speechbrain.nnet.CNN", Conv1d=_Conv1d)
    lin = mod("speechbrain.nnet.linear", Linear=O.SBLinear)
    emb = mod("speechbrain.nnet.embedding", Embedding=O.SBEmbedding)
    nnet = mod("speechbrain.nnet", CNN=cnn, linear=lin, embedding=emb)
    tr = mod("speechbrain.lobes.models.transformer.Transformer", PositionalEncoding=O.PositionalEncoding,
             TransformerEncoder=_TransformerEncoder, get_key_padding_mask=O.get_key_padding_mask,
             get_mask_from_lengths=O.get_mask_from_lengths)
    trp = mod("speechbrain.lobes.models.transformer", Transformer=tr)
    fs2 = mod("speechbrain.lobes.models.FastSpeech2", EncoderPreNet=O.EncoderPreNet,
              DurationPredictor=O.DurationPredictor, PostNet=O.PostNet, upsample=O.upsample,
              average_over_durations=O.average_over_durations, SSIMLoss=O.SSIMLoss)
    models = mod("speechbrain.lobes.models", transformer=trp, FastSpeech2=fs2)
    lobes = mod("speechbrain.lobes", models=models)
    sb = mod("speechbrain", nnet=nnet, lobes=lobes)
    return {m.__name__: m for m in (sb, nnet, cnn, lin, emb, lobes, models, trp, tr, fs2)}


def load():
    """-> (reference FastSpeech2 class, reference Loss class), executed from /root/reference with stubbed leaves."""
    assert available(), "the reference tree is not mounted"
    stubs = _stub_modules()
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        out = []
        for fname, alias in (("model.py", "_ref_fs2_model"), ("loss.py", "_ref_fs2_loss")):
            spec = importlib.util.spec_from_file_location(alias, os.path.join(REF_DIR, fname))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            out.append(m)
    finally:
        for k, v in saved.items():                       # speechbrain must not appear importable to anything else
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return out[0].FastSpeech2, out[1].Loss


def load_get_intensity_representation():
    """train.py:16-51 executed from its own source text (the module's top-level imports need speechbrain, tensorboard
    and the dataset stack, none of which the function uses)."""
    import ast
    import torch
    path = os.path.join(REF_DIR, "train.py")
    tree = ast.parse(open(path).read(), filename=path)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_intensity_representation")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["get_intensity_representation"]


def intensity_case(seed=11, B=5, Tp=23, D=5):
    """A collate-shaped 12-tuple (train.py:18-21) with ragged phoneme lengths, zero durations and padding, plus the frame
    intensities a stand-in extractor returns."""
    import torch
    g = torch.Generator().manual_seed(seed)
    phon_len = torch.tensor([23, 17, 9, 1, 12])
    dur = torch.randint(0, 7, (B, Tp), generator=g)
    dur[0, 3] = 0
    dur[3, 0] = 4
    for b in range(B):
        dur[b, int(phon_len[b]):] = 0
    mel_len = dur.sum(1)
    Tm = int(mel_len.max())
    frames = torch.randn(B, Tm, D, generator=g)
    phoneme = torch.randint(1, 90, (B, Tp), generator=g)
    rank_X = torch.randn(B, 82, Tm, generator=g)
    emo = torch.randint(0, 5, (B,), generator=g)
    batch = (phoneme, None, phon_len, None, None, None, dur, mel_len, None, None, rank_X, emo)
    return batch, frames


def run_reference_prototype_binning(entries_by_cell, speaker_list, emotion_list, bucket_size, feat_dim):
    """rank_model/inference.py:91-110 -- the `prototypes = np.zeros(...)` statement and the "Binning and averaging" loop
    of bucketize() -- executed from the reference's source on a caller-built `intensity_storage` (the part of the
    function before them is the dataloader / model loop).  The reference sizes the feature axis with n_emo; here it is
    passed as `n_emo` only for that statement via `feat_dim` when the two differ in a test."""
    import ast
    import warnings
    import numpy as np
    path = os.path.join(os.path.dirname(REF_DIR), "rank_model", "inference.py")
    tree = ast.parse(open(path).read(), filename=path)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "bucketize")
    alloc = next(n for n in fn.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "prototypes")
    loop = next(n for n in fn.body if isinstance(n, ast.For) and isinstance(n.iter, ast.Call)
                and "speaker_list" in ast.dump(n.iter) and "array_split" in ast.dump(n))
    ns = {"np": np, "intensity_storage": entries_by_cell, "speaker_list": speaker_list, "emotion_list": emotion_list,
          "n_spk": len(speaker_list), "n_emo": feat_dim, "bucket_size": bucket_size}
    exec(compile(ast.Module(body=[alloc], type_ignores=[]), path, "exec"), ns)
    assert ns["prototypes"].shape == (len(speaker_list), feat_dim, bucket_size, feat_dim)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                     # numpy: mean of an empty slice -> NaN, as in the reference
        exec(compile(ast.Module(body=[loop], type_ignores=[]), path, "exec"), ns)
    return ns["prototypes"]
