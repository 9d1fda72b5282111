"""The .npz utterance schema (rank_model/preprocess.py:134-151) and its reader (fastspeech2/dataset.py:12-58): write ->
read -> the reference-ordered collate, on the host.  When /root/reference is mounted the reader is also compared with the
reference's own FastSpeech2Dataset.__getitem__ executed from its source."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

PKG = "fine-grained-emotional-control-of-tts_b200"
REF = "/root/reference/emo_rank_tts/fastspeech2/dataset.py"


@pytest.fixture()
def files(tmp_path):
    io = importlib.import_module(PKG + ".npz_io")
    rng = np.random.default_rng(3)
    speakers, emotions = ["0011", "0012"], ["Angry", "Happy", "Neutral"]
    paths, raw = [], []
    for i in range(5):
        n_ph = int(rng.integers(4, 12))
        dur = rng.integers(0, 7, n_ph)
        dur[0] += 12
        T = int(dur.sum())
        phones = [io.DEFAULT_VALID_TOKENS[int(j)] for j in rng.integers(1, len(io.DEFAULT_VALID_TOKENS) - 1, n_ph)]
        rec = dict(phones=phones, emotion=emotions[i % 3], speaker=speakers[i % 2], audio_id=f"{i:06d}",
                   audio_path=f"/data/{i}.wav", transcript=f"hello {{noise}} world {i}", textgrid_path=f"/tg/{i}.TextGrid",
                   mel=rng.standard_normal((80, T)).astype(np.float32), pitch=rng.standard_normal(T),
                   energy=rng.standard_normal(T).astype(np.float32), durations=dur)
        p = str(tmp_path / f"{rec['emotion']}_{rec['audio_id']}.npz")
        io.write_npz_utterance(p, **rec)
        paths.append(p)
        raw.append(rec)
    (tmp_path / "fs2_train.txt").write_text("\n".join(paths) + "\n")
    return io, tmp_path, paths, raw, speakers, emotions


def test_schema_and_sample_dict(files):
    io, root, paths, raw, speakers, emotions = files
    assert sorted(np.load(paths[0], allow_pickle=True).files) == sorted(io.NPZ_KEYS)
    ds = io.NpzUtterances(str(root), "{noise}", speakers, emotions, mode="train")
    assert len(ds) == 5
    for i, rec in enumerate(raw):
        s = ds[i]
        assert s["mel"].dtype == torch.float32 and s["mel"].shape == (80, int(rec["durations"].sum()))
        assert torch.equal(s["duration"], torch.from_numpy(rec["durations"]).long())
        assert int(s["duration"].sum()) == s["mel"].shape[1] == len(s["pitch"]) == len(s["energy"])     # preprocess.py:133
        assert s["phoneme"].tolist() == [io.DEFAULT_VALID_TOKENS.index(p) for p in rec["phones"]]        # util.py:30-32
        assert int(s["speaker"]) == speakers.index(rec["speaker"]) and int(s["emotion"]) == emotions.index(rec["emotion"])
        assert s["text"] == f"hello  world {i}" and s["audio_path"] == rec["audio_path"]
        assert np.allclose(s["pitch"].numpy(), rec["pitch"].astype(np.float32))
    with pytest.raises(ValueError):
        io.phoneme2sequence(["NOT_A_PHONE"])
    other = str(root / "other.npz")
    np.savez(other, mel=np.zeros((80, 3)))
    with pytest.raises(KeyError):
        io.read_npz_utterance(other, speakers, emotions)


def test_samples_feed_the_collate(files):
    io, root, paths, raw, speakers, emotions = files
    data = importlib.import_module(PKG + ".data")
    ds = io.NpzUtterances(str(root), "{noise}", speakers, emotions)
    samples = [ds[i] for i in range(len(ds))]
    lens = torch.LongTensor([len(s["phoneme"]) for s in samples])
    order = torch.sort(lens, descending=True).indices.tolist()
    # the package's host collate takes (T, n_mels) mels and an intensity field; adapt and check the reference ordering
    utts = [dict(phoneme=s["phoneme"], duration=s["duration"], mel=s["mel"].t().contiguous(), pitch=s["pitch"],
                 energy=s["energy"], speaker=int(s["speaker"]), emotion=int(s["emotion"]),
                 intensity=torch.zeros(len(s["phoneme"]), 5)) for s in samples]
    batch, _ = data.collate(utts)
    assert batch[2].tolist() == [len(samples[j]["phoneme"]) for j in order]            # dataset.py:65-67
    assert batch[7].tolist() == [samples[j]["mel"].shape[1] for j in order]
    j0 = order[0]
    assert torch.equal(batch[3][0, :samples[j0]["mel"].shape[1]], samples[j0]["mel"].t())


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
def test_reader_matches_the_reference_dataset(files, monkeypatch):
    io, root, paths, raw, speakers, emotions = files
    util = types.ModuleType("util")
    util.phoneme2sequence = io.phoneme2sequence          # util.py needs speechbrain (absent): its table is restated in npz_io
    monkeypatch.setitem(sys.modules, "util", util)
    spec = importlib.util.spec_from_file_location("ref_fs2_dataset", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rds = ref.FastSpeech2Dataset(str(root), "{noise}", speakers, emotions, mode="train")
    mine = io.NpzUtterances(str(root), "{noise}", speakers, emotions, mode="train")
    assert len(rds) == len(mine)
    for i in range(len(mine)):
        a, b = rds[i], mine[i]
        assert a.keys() == b.keys()
        for k in a:
            if torch.is_tensor(a[k]):
                assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
            else:
                assert a[k] == b[k], k
