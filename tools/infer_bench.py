"""BASELINE configs[4]: FastSpeech2 batched inference with per-utterance emotion-intensity vectors, batch 256,
duration control (pace) 0.8 / 1.0 / 1.2 (reference call: inference.py:82 `model(phon_ids, spkr_ids, intensity=...)`).

Synthetic, random-init weights; the duration predictor's output bias is shifted to 1.7 (exp(1.7)-1 = 4.5 frames per
phoneme on average) because a random-init predictor says ~0 frames and there would be nothing to decode.  Inputs
arrive from pinned host memory and `mel_lens` / the mel itself are read back inside the timed region (e2e), as
the reference's inference loop hands the mel to the vocoder on the host side of the call.
Prints one JSON line per pace."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "fine-grained-emotional-control-of-tts_b200"


def fwd_flops(B, Tp, Tm, cfg):
    """Algorithmic dense FLOPs of one forward on the padded rectangle."""
    D, F, H = cfg["enc_d_model"], cfg["enc_ffn_dim"], cfg["enc_num_head"]
    k0, k1 = cfg["ffn_cnn_kernel_size_list"]
    layer = lambda T: B * T * (2 * D * 3 * D + 2 * D * D + 2 * D * F * k0 + 2 * F * D * k1) + B * H * (2 * T * T * (D // H)) * 2
    f = cfg["enc_num_layers"] * layer(Tp) + cfg["dec_num_layers"] * layer(Tm)
    f += 3 * B * Tp * (2 * 2 * D * D * cfg["dur_pred_kernel_size"])
    E, kp, nm = cfg["postnet_embedding_dim"], cfg["postnet_kernel_size"], cfg["n_mels"]
    f += B * Tm * 2 * kp * (nm * E + (cfg["postnet_n_convolutions"] - 2) * E * E + E * nm)
    return f + B * Tp * 2 * D * (2 * D + 5) + B * Tm * 2 * D * nm


def main():
    pkg = importlib.import_module(PKG)
    B = int(os.environ.get("INFER_BATCH", "256"))
    steps = int(os.environ.get("INFER_STEPS", "10"))
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision="bf16").to(dev).eval()
    with torch.no_grad():
        model.state_dict()["durPred.linear.w.bias"].fill_(1.7)
    g = torch.Generator().manual_seed(1234)
    R = 4
    batches = []
    for r in range(R):
        lens = torch.randint(24, 129, (B,), generator=g).sort(descending=True).values      # dataset.py:65-67 order
        Tp = int(lens[0])
        tokens = torch.randint(1, 95, (B, Tp), generator=g)
        tokens[torch.arange(Tp)[None] >= lens[:, None]] = 0
        speakers = torch.randint(0, 4, (B,), generator=g)
        proto = torch.randn(B, 1, 5, generator=g)                                          # inference.py:17-19
        intensity = (proto.expand(B, Tp, 5) * (tokens != 0).unsqueeze(-1)).contiguous()
        batches.append(tuple(t.pin_memory() for t in (tokens, speakers, intensity)))

    def run(batch, pace):
        tokens, speakers, intensity = (t.to(dev, non_blocking=True) for t in batch)
        with torch.no_grad():
            out = model(tokens, speakers, pace=pace, intensity=intensity)
        mel_host = out[0].to("cpu", non_blocking=True)        # (B, Tm, 80) -> vocoder side
        return out, mel_host

    rows = []
    for pace in (0.8, 1.0, 1.2):
        for i in range(2 * R):                      # every batch shape twice: the workspace arenas reach their final size
            run(batches[i % R], pace)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        frames = 0
        d2h = 0
        flops = 0.0
        e0.record()
        for i in range(steps):
            out, mel_host = run(batches[i % R], pace)
            frames += int(out[7].sum())
            d2h += mel_host.numel() * 4
            flops += fwd_flops(B, batches[i % R][0].shape[1], int(out[0].shape[1]), pkg.DEFAULT_MODEL_CONFIG)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rows.append({
            "metric": "fastspeech2_inference_mel_frames_per_sec", "value": frames / (ms * 1e-3), "unit": "mel_frames/s",
            "n_gpus": 1, "steps": steps, "ms_per_step": ms / steps, "dtype": "bf16", "data": "synthetic",
            "tflops": flops / (ms * 1e-3) / 1e12,
            "config": {"workload": "FastSpeech2 batched inference (BASELINE configs[4])", "batch": B, "pace": pace,
                       "phonemes": "U{24..128} per utterance", "valid_frames_per_step": frames // steps,
                       "padded_Tm_last": int(out[0].shape[1]),
                       "intensity": "per-utterance (5,) prototype broadcast over Tp",
                       "e2e": "pinned host inputs -> device, mel (B,Tm,80) fp32 + mel_lens -> host inside the timed region",
                       "d2h_bytes_per_step": d2h // steps}})
        print(json.dumps(rows[-1]), flush=True)
    return rows


if __name__ == "__main__":
    main()
