"""Parity at the BENCHMARK configurations (BASELINE.json configs[1] t5-base and configs[3] t5-large shapes), on the GPU.

The oracle (oracle/vid2seq_oracle.py, pinned to the real reference by tests/test_oracle_cpu.py + tests/golden) runs on
the same B200 in three arithmetic modes:
  fp32                         = the reference's numerics (tier B reference),
  emu  (bf16 operands, fp32 accumulate, the kernels' rounding points)   = tier A reference,
  emu64 (same bf16 operands, products accumulated in fp64, rounded once) = the accumulation-order-free version of emu.
|emu - emu64| is the MEASURED noise floor of tier A: the distance between two legal evaluations of the same
bf16-operand arithmetic (a 1-ulp fp32 difference flips bf16 roundings downstream and the flips amplify through 36
residual sub-layers).  Gates, no escape hatches:
  * end to end: rel-L2(cuda, emu64) <= max(1e-3, 1.5 x rel-L2(emu, emu64));
  * every time-token argmax mismatch is a tie inside the oracle's own noise band;
  * teacher-forced: each CUDA sub-layer fed the ORACLE's input reproduces the oracle's output to <= 1e-3 (north-star
    tolerance; measured ~1e-4), which proves the end-to-end figure is amplification and not a kernel error;
  * gradients: per-parameter rel-L2 vs emu at <= 1.5 x the emu/emu64 floor (worst and median).
Every measured number is written to gpurun_out/r02_parity.json (committed copy: profiles/r02_parity.json).
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RESULTS = {}


class Tok:
    pad_token_id, eos_token_id = 0, 1

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


@pytest.fixture(scope="module", autouse=True)
def _dump_results():
    yield
    if not RESULTS:
        return
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        path = os.path.join(out_dir, "r02_parity.json")
        old = {}
        if os.path.isfile(path):
            try:
                old = json.load(open(path))
            except Exception:
                old = {}
        old.update(RESULTS)
        with open(path, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)
    except OSError:
        pass


def record(key, **vals):
    RESULTS.setdefault(key, {}).update({k: (float(v) if isinstance(v, (int, float)) else v) for k, v in vals.items()})
    print(f"[parity] {key}: " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in vals.items()))


def build(cfg, name):
    from vidchapters_b200 import Vid2Seq
    tok = Tok(cfg["base_vocab"] + cfg["num_bins"])
    m = Vid2Seq(name, num_features=cfg["num_features"], embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                heads=cfg["heads"], mlp_dim=cfg["mlp_dim"], vis_drop=0.0, tokenizer=tok, enc_drop=0.0, dec_drop=0.0,
                num_bins=cfg["num_bins"], t5_config=cfg, seed=0)
    return m.to("cuda")


def batch(B, T, L, S, seed, cfg):
    from bench import synth_batch
    v, i, o = synth_batch(B, T, L, S, seed, cfg["base_vocab"], cfg["base_vocab"] + cfg["num_bins"])
    return v.cuda(), i.cuda(), o.cuda()


def _cfg(name):
    from vidchapters_b200 import T5_BASE, T5_LARGE
    return dict(T5_BASE if name == "t5-base" else T5_LARGE)


def _oracle(sd, cfg, video, inp, out, **kw):
    from oracle import vid2seq_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    return O.vid2seq_forward(sd, cfg, video, inp, inp != 0, out, out != 0, **kw)


# B = the per-GPU batch BASELINE.json quotes for the config (16 for t5-base, 8 for t5-large)
@pytest.mark.parametrize("name,B", [("t5-base", 16), ("t5-large", 8)])
def test_full_config_logits_loss_argmax_vs_oracle(name, B):
    cfg = _cfg(name)
    T, L, S = 100, 1000, 256
    m = build(cfg, name)
    video, inp, out = batch(B, T, L, S, 1234, cfg)
    tok = lambda x: {"input_ids": x, "attention_mask": x != 0}
    loss, logits = m.forward_logits(video, tok(inp), tok(out))
    sd = {k: v.detach() for k, v in m._params.items()}
    V0 = cfg["base_vocab"]
    valid = (out != 0)
    with torch.no_grad():
        o_emu = _oracle(sd, cfg, video, inp, out, emulate_bf16=True, flash_rounding=True)
        l_emu, z_emu = o_emu["loss"].item(), o_emu["logits"]
        del o_emu
        o64 = _oracle(sd, cfg, video, inp, out, emulate_bf16=True, flash_rounding=True, acc64=True)
        l_64, z_64 = o64["loss"].item(), o64["logits"]
        del o64
        o32 = _oracle(sd, cfg, video, inp, out)
        l_32, z_32 = o32["loss"].item(), o32["logits"]
        del o32
    # rows of padded target positions never reach the loss, but they are real decoder outputs: compare everything
    floor = rel(z_emu, z_64)
    e_cuda_64 = rel(logits, z_64)
    e_cuda_emu = rel(logits, z_emu)
    e_cuda_32 = rel(logits, z_32)
    e_emu_32 = rel(z_emu, z_32)
    record(f"{name}.B{B}.logits", floor_emu_vs_emu64=floor, cuda_vs_emu64=e_cuda_64, cuda_vs_emu=e_cuda_emu,
           cuda_vs_fp32=e_cuda_32, emu_vs_fp32=e_emu_32, loss_cuda=loss.item(), loss_emu=l_emu, loss_emu64=l_64,
           loss_fp32=l_32)
    assert e_cuda_64 <= max(1e-3, 1.5 * floor), (e_cuda_64, floor)
    # tier B: against the fp32 reference numerics, no worse than 1.5 x what the bf16-operand arithmetic itself costs
    assert e_cuda_32 <= 1.5 * e_emu_32 + 1e-3, (e_cuda_32, e_emu_32)
    assert abs(loss.item() - l_64) <= max(5e-4, 3 * abs(l_emu - l_64) / abs(l_64)) * abs(l_64)
    assert abs(loss.item() - l_32) <= 2e-3 * abs(l_32)
    # ---- time-token argmax: bit-exact except ties inside the oracle's own noise band
    tz_c, tz_e, tz_6 = logits[..., V0:], z_emu[..., V0:], z_64[..., V0:]
    noise = (tz_e - tz_6).abs().max().item()               # what two legal evaluations of the arithmetic differ by
    a_c, a_6 = tz_c.argmax(-1), tz_6.argmax(-1)
    top2 = tz_6.topk(2, dim=-1).values
    gap = (top2[..., 0] - top2[..., 1])
    mism = (a_c != a_6)
    mism_valid = mism & valid
    n_mism = int(mism.sum())
    max_gap = gap[mism].max().item() if n_mism else 0.0
    oracle_self = int((tz_e.argmax(-1) != a_6).sum())
    record(f"{name}.B{B}.time_argmax", positions=int(a_c.numel()), mismatches=n_mism,
           mismatches_on_valid_targets=int(mism_valid.sum()), max_oracle_gap_at_mismatch=max_gap,
           oracle_noise_max_abs=noise, emu_vs_emu64_mismatches=oracle_self)
    assert max_gap <= 3.0 * noise, (max_gap, noise)        # = 1.5 x (both candidates moving by the noise)


def _grad_stats(ga, gb, layout):
    errs = []
    for n, (o, shp, k) in layout.items():
        errs.append((rel(ga[o:o + k], gb[o:o + k]), n))
    errs.sort(reverse=True)
    return errs


@pytest.mark.parametrize("name,B", [("t5-base", 4), ("t5-large", 2)])
def test_full_config_gradients_vs_oracle(name, B):
    """Same sequence lengths as the benchmark (100 / 1000 / 256), smaller batch so that the autograd oracle (which keeps
    every (B,H,L,L) tensor) runs three times in memory."""
    cfg = _cfg(name)
    T, L, S = 100, 1000, 256
    m = build(cfg, name)
    m.train()
    video, inp, out = batch(B, T, L, S, 4321, cfg)
    tok = lambda x: {"input_ids": x, "attention_mask": x != 0}
    ld, _ = m(video, tok(inp), tok(out))
    ld["loss"].backward()
    eng = m.engine
    g_cuda = eng.flat_g.clone()

    def oracle_grads(**kw):
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in m._params.items()}
        o = _oracle(sd, cfg, video, inp, out, **kw)
        o["loss"].backward()
        flat = torch.zeros_like(g_cuda)
        for n, (off, shp, k) in eng.layout.items():
            flat[off:off + k] = sd[n].grad.reshape(-1)
        return o["loss"].item(), flat

    l_emu, g_emu = oracle_grads(emulate_bf16=True, flash_rounding=True)
    l_64, g_64 = oracle_grads(emulate_bf16=True, flash_rounding=True, acc64=True)
    e_floor = _grad_stats(g_emu, g_64, eng.layout)
    e_cuda = _grad_stats(g_cuda, g_64, eng.layout)
    med = lambda e: e[len(e) // 2][0]
    record(f"{name}.B{B}.grads", n_params=len(e_cuda), cuda_vs_emu64_worst=e_cuda[0][0], cuda_worst_name=e_cuda[0][1],
           cuda_vs_emu64_median=med(e_cuda), floor_worst=e_floor[0][0], floor_worst_name=e_floor[0][1],
           floor_median=med(e_floor), flat_cuda_vs_emu64=rel(g_cuda, g_64), flat_floor=rel(g_emu, g_64),
           loss_cuda=ld["loss"].item(), loss_emu=l_emu, loss_emu64=l_64)
    assert med(e_cuda) <= max(1e-3, 1.5 * med(e_floor)), (med(e_cuda), med(e_floor))
    assert e_cuda[0][0] <= max(1e-3, 1.5 * e_floor[0][0]), (e_cuda[:3], e_floor[:3])
    assert rel(g_cuda, g_64) <= max(1e-3, 1.5 * rel(g_emu, g_64))


@pytest.mark.parametrize("name,B", [("t5-base", 2), ("t5-large", 1)])
def test_teacher_forced_sublayers(name, B):
    """Every CUDA residual sub-layer (and both final norms and the LM head) fed the ORACLE's own input at the benchmark
    sequence lengths: the output must match the oracle's to the north-star tolerance 1e-3 (expected ~1e-4: the only
    differences left are the fp32 summation order inside one sub-layer)."""
    from parity_util import teacher_forced_errors
    cfg = _cfg(name)
    m = build(cfg, name)
    video, inp, out = batch(B, 100, 1000, 256, 99, cfg)
    errs = teacher_forced_errors(m, cfg, video, inp, out)
    worst_branch = max(errs, key=lambda t: t[1])
    record(f"{name}.B{B}.teacher_forced", sublayers=len(errs), worst_out=errs[0][0], worst_out_name=errs[0][2],
           median_out=errs[len(errs) // 2][0], worst_branch=worst_branch[1], worst_branch_name=worst_branch[2])
    assert errs[0][0] <= 1e-3, errs[:5]
