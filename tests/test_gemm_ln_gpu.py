"""fs2_gemm_ln_tc (GEMM + bias + branch dropout + residual + LayerNorm in one tcgen05 kernel, gemm_ln.cu) against

  * the two-launch form it replaces (fs2_gemm_tc with an fp32 branch in HBM, then fs2_ln_fwd): same accumulators, same
    dropout mask, same statistics formula -- outputs agree to fp32 rounding, reflect-halo / zero rows included;
  * a float64 torch restatement of  LN(x + Linear(a))  (reference: speechbrain's post-norm TransformerEncoderLayer reached
    from fastspeech2/model.py:344-347, 425-428), dropout off;

and fs2_ln_bwd's `y` form (x_hat recovered from the forward OUTPUT, for forwards that never store the branch) against its
ordinary form on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
PAD = 4
C = 384


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def _struct(cls, **kw):
    p = cls()
    for k, v in kw.items():
        setattr(p, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    return p


def _case(lib, B, T, K, halo, p_drop, seed=0):
    rows = B * (T + 2 * PAD)
    A = _rand((rows, K), seed + 1, 0.5).to(torch.bfloat16)
    W = _rand((C, K), seed + 2, K ** -0.5).to(torch.bfloat16)
    bias = _rand((C,), seed + 3, 0.1)
    x = _rand((rows, C), seed + 4)
    gamma = _rand((C,), seed + 5, 0.2) + 1
    beta = _rand((C,), seed + 6, 0.1)
    ctr = torch.tensor([seed + 17], dtype=torch.int64, device="cuda")       # device-side dropout step counter
    eps, dseed = 1e-6, 0xABCDEF12345
    # ---- two launches: branch in HBM
    proj = torch.empty(rows, C, device="cuda")
    lib.gemm(mode=0, M=rows, N=C, K=K, A=A, lda=K, a_rows=rows, a_inner=K, B=W, ldb=K, b_rows=C, b_inner=K, Cout=proj,
             ldc=C, c_bf16=False, ab_bf16=True, bias=bias, rs_T=T, rs_Tp=T + 2 * PAD)
    ref_f = torch.full((rows, C), 7.0, device="cuda")
    ref_a = torch.full((rows, C), 7.0, device="cuda", dtype=torch.bfloat16)
    ref_mean, ref_rstd = torch.zeros(rows, device="cuda"), torch.zeros(rows, device="cuda")
    p = _struct(lib.Fs2LnFwd, B=B, T=T, C=C, x=x, branch=proj, drop_b_p=p_drop, drop_b_seed=dseed, gamma=gamma, beta=beta,
                eps=eps, out_f32=ref_f, out_act=ref_a, act_bf16=1, halo=halo, mean=ref_mean, rstd=ref_rstd, seed_dev=ctr)
    lib.call("fs2_ln_fwd", lib.C.addressof(p))
    # ---- one launch
    out_f = torch.full((rows, C), 7.0, device="cuda")
    out_a = torch.full((rows, C), 7.0, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.zeros(rows, device="cuda"), torch.zeros(rows, device="cuda")
    q = _struct(lib.Fs2GemmLn, B=B, T=T, K=K, lda=K, ldw=K, A=A, W=W, bias=bias, x=x, drop_p=p_drop, drop_seed=dseed,
                seed_dev=ctr, gamma=gamma, beta=beta, eps=eps, out_f32=out_f, out_act=out_a, halo=halo, mean=mean, rstd=rstd)
    lib.call("fs2_gemm_ln_tc", lib.C.addressof(q))
    torch.cuda.synchronize()
    assert lib.gemm_tc_error_flag() == 0, "tcgen05 kernel reported an mbarrier timeout"
    return dict(A=A, W=W, bias=bias, x=x, gamma=gamma, beta=beta, eps=eps, proj=proj, ctr=ctr, dseed=dseed,
                ref_f=ref_f, ref_a=ref_a, ref_mean=ref_mean, ref_rstd=ref_rstd, out_f=out_f, out_a=out_a, mean=mean,
                rstd=rstd, rows=rows)


CASES = [
    # B, T, K, halo, dropout
    (3, 37, 384, 4, 0.0),          # several items inside one 128-row tile, mirrors at both ends
    (5, 9, 384, 4, 0.0),           # short items: a row can own both mirrors; M not a multiple of 128
    (2, 150, 1536, 0, 0.2),        # FFN conv 2 shape (24 k-blocks: the ring wraps 8 times), halo rows zeroed, dropout
    (4, 131, 384, 4, 0.1),
    (32, 800, 384, 4, 0.1),        # BASELINE configs[2] decoder rows: 202 tiles on 148 SMs (two tiles per CTA)
    (32, 120, 1536, 0, 0.1),       # encoder rows
]


@pytest.mark.parametrize("B,T,K,halo,p_drop", CASES)
def test_fused_gemm_layernorm_equals_two_launches(lib, B, T, K, halo, p_drop):
    r = _case(lib, B, T, K, halo, p_drop)
    # fp32 output incl. mirror / zero / untouched (sentinel 7.0) rows: only the summation order of the statistics differs
    assert (r["out_f"] - r["ref_f"]).abs().max().item() <= 2e-5
    rect = torch.zeros(B, T + 2 * PAD, dtype=torch.bool, device="cuda")
    rect[:, PAD:PAD + T] = True
    rect = rect.view(-1)
    assert (r["mean"] - r["ref_mean"])[rect].abs().max().item() <= 1e-6
    assert ((r["rstd"] - r["ref_rstd"])[rect].abs() / r["ref_rstd"][rect]).max().item() <= 1e-5
    assert torch.equal(r["mean"][~rect], torch.zeros_like(r["mean"][~rect]))     # statistics of halo rows are not written
    # bf16 operand copy: roundings of values that agree to 2e-5 -- at most one bf16 step apart, almost all identical
    a, b = r["out_a"].float(), r["ref_a"].float()
    assert ((a - b).abs() <= 2.0 ** -7 * b.abs().clamp_min(2.0 ** -6)).all()
    assert (a == b).float().mean().item() >= 0.995


@pytest.mark.parametrize("B,T,K,halo", [(3, 37, 384, 4), (2, 150, 1536, 0)])
def test_fused_gemm_layernorm_matches_float64_torch(lib, B, T, K, halo):
    r = _case(lib, B, T, K, halo, 0.0)
    z = r["x"].double() + r["A"].double() @ r["W"].double().t() + r["bias"].double()
    ref = F.layer_norm(z, (C,), r["gamma"].double(), r["beta"].double(), r["eps"])
    got = r["out_f"].view(B, T + 2 * PAD, C)[:, PAD:PAD + T].double()
    # bf16 products are exact in fp32; what is left is fp32 accumulation over K and the fp32 statistics: 1e-4 absolute
    assert (got - ref.view(B, T + 2 * PAD, C)[:, PAD:PAD + T]).abs().max().item() <= 1e-4
    # reflect halo of the bf16 copy (feeds the k = 9 conv): row -i mirrors row i, row T-1+i mirrors row T-1-i
    oa = r["out_a"].view(B, T + 2 * PAD, C)
    for i in range(1, halo + 1):
        assert torch.equal(oa[:, PAD - i], oa[:, PAD + i])
        assert torch.equal(oa[:, PAD + T - 1 + i], oa[:, PAD + T - 1 - i])
    if halo == 0:
        assert oa[:, :PAD].abs().sum() == 0 and oa[:, PAD + T:].abs().sum() == 0


def test_fused_gemm_layernorm_dropout_rate_and_scaling(lib):
    """identity LayerNorm statistics aside, the kept branch elements are scaled by 1/(1-p): checked through the row means"""
    B, T, K, p_drop = 8, 200, 384, 0.3
    r = _case(lib, B, T, K, 0, p_drop, seed=5)
    # mean of z = mean(x) + mean(keep * branch): against the two-launch form it is already exact (previous test); here
    # the mask itself: reconstruct z from the outputs and compare the dropped fraction of branch elements
    rect = torch.zeros(B, T + 2 * PAD, dtype=torch.bool, device="cuda")
    rect[:, PAD:PAD + T] = True
    rect = rect.view(-1)
    xhat = (r["out_f"] - r["beta"]) / r["gamma"]
    z = xhat / r["rstd"][:, None] + r["mean"][:, None]
    kept_branch = (z - r["x"])[rect]
    branch = r["proj"][rect]
    dropped = kept_branch.abs() < 1e-3 * branch.abs().clamp_min(1e-2)
    big = branch.abs() > 0.05
    frac = dropped[big].float().mean().item()
    assert abs(frac - p_drop) < 0.01
    keep = ~dropped & big
    assert torch.allclose(kept_branch[keep], branch[keep] / (1 - p_drop), atol=2e-3, rtol=2e-3)


@pytest.mark.parametrize("B,T,K,halo,p_drop", [(3, 37, 384, 4, 0.0), (4, 131, 1536, 0, 0.25)])
def test_layernorm_backward_from_the_forward_output(lib, B, T, K, halo, p_drop):
    r = _case(lib, B, T, K, halo, p_drop, seed=3)
    rows = r["rows"]
    dy = _rand((rows, C), 11)
    dy2 = _rand((rows, C), 12)

    def run(use_y):
        dx = torch.full((rows, C), float("nan"), device="cuda")
        dact = torch.full((rows, C), 7.0, device="cuda", dtype=torch.bfloat16)
        dg, db, cs = (torch.zeros(C, device="cuda") for _ in range(3))
        kw = dict(B=B, T=T, C=C, dy=dy, dy2=dy2, dy2_fold=halo, gamma=r["gamma"], beta=r["beta"], eps=r["eps"],
                  drop_b_p=p_drop, drop_b_seed=r["dseed"], rstd=r["rstd"], dx_f32=dx, dact=dact, act_bf16=1, dgamma=dg,
                  dbeta=db, dact_colsum=cs, seed_dev=r["ctr"])
        if use_y:
            kw.update(y=r["out_f"])
        else:
            kw.update(x=r["x"], branch=r["proj"], mean=r["mean"])
        q = _struct(lib.Fs2LnBwd, **kw)
        lib.call("fs2_ln_bwd", lib.C.addressof(q))
        torch.cuda.synchronize()
        return dx, dact.float(), dg, db, cs

    a, b = run(False), run(True)
    for name, u, v in zip(("dx", "dact", "dgamma", "dbeta", "dact_colsum"), a, b):
        tol = 2e-4 * max(1.0, u.abs().max().item())
        if name == "dact":
            tol = 1e-2 * max(1.0, u.abs().max().item())          # bf16 roundings of values that agree to 2e-4
        if name == "dact_colsum":
            tol = 1e-3 * max(1.0, u.abs().max().item())          # fp32 sums over all rows
        assert (u - v).abs().max().item() <= tol, name
