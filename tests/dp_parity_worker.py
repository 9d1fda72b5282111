"""Data-parallel parity on hardware -- run by tests/test_dp_nccl_gpu.py as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dp_parity_worker.py

One process per GPU over NCCL, the bench's own configuration (CUDA-graph replay, async mel_lens, early all-reduce of the
decoder piece through `grad_ready_hook`, per-piece AdamW):

  phase 1 (lr = 0, so parameters stay put): three DataParallelStep calls on the same per-rank batch -- eager, capture,
           replay -- and after each one   flat_grad / world  ==  mean over ranks of each rank's fp64-oracle gradient;
  phase 2 (lr = 1e-4): three more steps on changing batches, then the replicas' flat parameter buffers must be
           bit-identical across ranks and must have moved.

Each rank prints one JSON line `DPRESULT {...}`; the parent asserts on them."""
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fs2_oracle as O  # noqa: E402

PKG = "fine-grained-emotional-control-of-tts_b200"


def main():
    pkg = importlib.import_module(PKG)
    par = importlib.import_module(PKG + ".parallel")
    data = importlib.import_module(PKG + ".data")
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    rank, world, local = par.init_from_env("nccl")
    assert world > 1, "launch under torchrun with >= 2 ranks"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B = int(os.environ.get("DP_BATCH", "32"))
    batches = data.synthetic_batches(B, 4, seed=1234, rank=rank, world=world)         # this rank's shard, bench recipe
    oracle = O.build(seed=0, dtype=torch.float64).to(dev).eval()
    keys = [k for k, _ in oracle.named_parameters()]

    def oracle_mean_grad(batch, intensity):
        tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = [t.to(dev) for t in batch[:8]]
        oracle.zero_grad(set_to_none=True)
        d = lambda t: t.double()
        preds = oracle(tokens, speakers, dur, d(pitch), d(energy), intensity=d(intensity.to(dev)))
        loss = O.Loss(**O.DEFAULT_LOSS_CONFIG)(preds, (d(mel), dur, d(pitch), d(energy), out_lens, in_lens), 0)
        loss["total_loss"].backward()
        grads = {k: p.grad.detach().clone() for k, p in oracle.named_parameters()}
        flat = torch.cat([grads[k].flatten() for k in keys])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= world
        out, o = {}, 0
        for k in keys:
            n = grads[k].numel()
            out[k] = flat[o:o + n].view_as(grads[k])
            o += n
        oracle.zero_grad(set_to_none=True)
        return out, flat

    result = {"rank": rank, "world": world, "B": B}
    for precision in ("fp32", "bf16"):
        model = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision=precision)
        model.load_state_dict({k: v.detach().float().cpu() for k, v in oracle.state_dict().items()})
        model = model.to(dev).eval()                              # dropout off: the oracle comparison needs it
        model.use_cuda_graphs = True
        model.async_mel_lens = True
        crit = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)
        opt = pkg.FusedAdamW(model, lr=0.0, weight_decay=0.0)
        trainer = par.DataParallelStep(model, crit, opt)
        batch, intensity = batches[0]
        gref, gref_flat = oracle_mean_grad(batch, intensity)
        dbatch = [t.to(dev) for t in batch[:8]]
        dint = intensity.to(dev)
        p0 = model.store.flat.clone()
        phase1 = []
        for step in range(3):                                    # eager, capture, replay
            trainer(dbatch, dint)
            torch.cuda.synchronize()
            mine = {k: p.grad.double() / world for k, p in model.named_parameters()}
            flat = torch.cat([mine[k].flatten() for k in keys])
            rl2 = ((flat - gref_flat).norm() / gref_flat.norm()).item()
            worst = max(((mine[k] - gref[k]).abs().max() / (gref[k].abs().max() + 1e-30)).item() for k in keys)
            phase1.append({"rl2": rl2, "worst_maxnorm": worst})
        graphs = len(model._graphs)
        unchanged = bool(torch.equal(p0, model.store.flat))     # lr = 0
        # phase 2: real updates, replicas must stay bit-identical
        opt.lr, opt.weight_decay = 1e-4, 1e-2
        for step in range(3):
            b, it = batches[1 + step % 3]
            trainer([t.to(dev) for t in b[:8]], it.to(dev))
        torch.cuda.synchronize()
        flat = model.store.flat
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        identical = all(bool(torch.equal(gathered[0], g)) for g in gathered[1:])
        moved = float((flat - p0).abs().max())
        result[precision] = {"phase1": phase1, "graphs_captured": graphs, "params_unchanged_at_lr0": unchanged,
                             "replicas_bit_identical": identical, "max_param_move": moved,
                             "tc_error_flag": int(importlib.import_module(PKG + "._lib").gemm_tc_error_flag())}
        del model, trainer, opt
        torch.cuda.empty_cache()
    print("DPRESULT " + json.dumps(result), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
