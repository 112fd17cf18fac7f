"""More end-to-end GPU parity: the vc.py path (no time tokens, ragged vocabulary, `padding="longest"` batches, int64
masks), the default two-pass dvc.py step (generative + denoising pass sharing the visual-encoder output) against the
oracle, and the multi-GPU data-parallel check (self-skips on a single GPU)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


class Tok:
    pad_token_id, eos_token_id = 0, 1

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def build(cfg, seed=0):
    from vidchapters_b200 import Vid2Seq
    tok = Tok(cfg["base_vocab"] + cfg["num_bins"])
    m = Vid2Seq("t5-base", num_features=cfg["num_features"], embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                heads=cfg["heads"], mlp_dim=cfg["mlp_dim"], vis_drop=0.0, tokenizer=tok, enc_drop=0.0, dec_drop=0.0,
                num_bins=cfg["num_bins"], t5_config=cfg, seed=seed)
    return m.to("cuda")


def _floor_gated_grads(m, sd, sd64, tag):
    errs, ferrs = [], []
    for n, p in m._params.items():
        if sd64[n].grad is None:
            continue
        errs.append((rel(p.grad, sd64[n].grad), n))
        ferrs.append((rel(sd[n].grad, sd64[n].grad), n))
    errs.sort(reverse=True)
    ferrs.sort(reverse=True)
    med, fmed = errs[len(errs) // 2][0], ferrs[len(ferrs) // 2][0]
    print(f"[{tag}] gradients vs emu64: worst {errs[0][0]:.3e} ({errs[0][1]}), median {med:.3e}; floor worst "
          f"{ferrs[0][0]:.3e}, median {fmed:.3e}")
    assert med <= max(1e-3, 2.0 * fmed), (med, fmed)           # (2-layer models: see tests/test_e2e_gpu.py)
    assert errs[0][0] <= max(1e-3, 2.0 * ferrs[0][0]), (errs[:3], ferrs[:3])


def test_vc_path_golden():
    """vc.py:26-86,280,299-312: num_bins=0 (no time tokens, no renorm), vocabulary 1012 (not a multiple of 8, like
    32100), `padding="longest"` ragged batch (37 / 19 tokens), clips of 7 < max_feats frames, int64 attention masks as
    the HF tokenizer returns them, stock clip_grad_norm_ + torch Adam as vc.py runs them."""
    from oracle import vid2seq_oracle as O
    fx = torch.load(os.path.join(GOLD, "tiny_vc.pt"), weights_only=False)
    cfg = fx["cfg"]
    m = build(cfg)
    m.train()
    video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
    it = {"input_ids": inp, "attention_mask": (inp != 0).long()}
    ot = {"input_ids": out, "attention_mask": (out != 0).long()}
    loss, logits = m.forward_logits(video, it, ot)
    assert logits.shape[-1] == 1012
    assert abs(loss.item() - fx["loss"].item()) < 2e-3 * abs(fx["loss"].item())
    eb = rel(logits.cpu(), fx["logits"])
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m._params.items()}
    sd64 = {k: v.detach().clone().requires_grad_(True) for k, v in m._params.items()}
    o = O.vid2seq_forward(sd, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True)
    o64 = O.vid2seq_forward(sd64, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True,
                            acc64=True)
    floor, ea = rel(o["logits"], o64["logits"]), rel(logits, o64["logits"])
    print(f"[vc path] logits vs fp32 reference {eb:.3e}; vs emu64 {ea:.3e} (floor {floor:.3e})")
    assert eb < 1.2e-2 and ea <= max(1e-3, 1.5 * floor)
    ld, _ = m(video, it, ot)
    ld["loss"].backward()
    o["loss"].backward()
    o64["loss"].backward()
    _floor_gated_grads(m, sd, sd64, "vc path")
    for n, gn in fx["grad_norms"].items():
        assert abs(m._params[n].grad.norm().item() - gn) <= 6e-2 * gn + 1e-9, n
    opt = torch.optim.Adam(m.parameters(), lr=3e-4)
    torch.nn.utils.clip_grad_norm_(m.parameters(), 0.1)
    opt.step()
    assert rel(m._params["t5_model.encoder.final_layer_norm.weight"].cpu(), fx["after_step"]["enc_ln"]) < 1e-4
    l2, _ = m(video, it, ot)
    assert l2["loss"].item() < ld["loss"].item()


def test_two_pass_step_matches_oracle():
    """dvc.py:70-100 (the reference's default step): generative pass + denoising pass that consumes the cached
    `video_dict`; both losses are summed before one backward, so the visual encoder receives gradient from both."""
    from oracle import vid2seq_oracle as O
    fx = torch.load(os.path.join(GOLD, "tiny_long.pt"), weights_only=False)
    cfg = fx["cfg"]
    m = build(cfg)
    m.train()
    video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
    g = torch.Generator().manual_seed(11)
    V = cfg["base_vocab"] + cfg["num_bins"]
    B = video.shape[0]
    inp2 = torch.randint(2, V, (B, 230), generator=g); inp2[1, -40:] = 0        # denoising input (~0.75 L + sentinels)
    out2 = torch.randint(2, V, (B, 90), generator=g); out2[0, -11:] = 0         # denoising target
    inp2, out2 = inp2.cuda(), out2.cuda()
    tok = lambda x: {"input_ids": x, "attention_mask": x != 0}
    l1, vd = m(video, tok(inp), tok(out))
    l2, _ = m(vd, tok(inp2), tok(out2))
    (l1["loss"] + l2["loss"]).backward()

    def oracle(**kw):
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in m._params.items()}
        o1 = O.vid2seq_forward(sd, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True, **kw)
        o2 = O.vid2seq_forward(sd, cfg, o1["video"], inp2, inp2 != 0, out2, out2 != 0, emulate_bf16=True,
                               flash_rounding=True, video_is_cached=True, **kw)
        (o1["loss"] + o2["loss"]).backward()
        return o1["loss"].item(), o2["loss"].item(), sd

    a1, a2, sd = oracle()
    b1, b2, sd64 = oracle(acc64=True)
    print(f"[two-pass] losses cuda {l1['loss'].item():.5f} + {l2['loss'].item():.5f}; emu64 {b1:.5f} + {b2:.5f}")
    assert abs(l1["loss"].item() - b1) <= max(5e-4, 3 * abs(a1 - b1) / b1) * b1
    assert abs(l2["loss"].item() - b2) <= max(5e-4, 3 * abs(a2 - b2) / b2) * b2
    _floor_gated_grads(m, sd, sd64, "two-pass")
    # the visual encoder really saw both losses: its gradient differs from the one-pass gradient
    g_two = m._params["visual_encoder.blocks.0.attn.qkv.weight"].grad.clone()
    for p in m.parameters():
        p.grad = None
    l1b, _ = m(video, tok(inp), tok(out))
    l1b["loss"].backward()
    assert rel(m._params["visual_encoder.blocks.0.attn.qkv.weight"].grad, g_two) > 0.05


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_nccl_overlapped_equals_flat_allreduce():
    """tools/dp_check_gpu.py under torchrun on 2 GPUs: the graphed step with the region-wise all-reduce overlapped with
    the backward trains exactly like the eager step with one flat all-reduce, and all ranks end bit-identical."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(ROOT, "tools", "dp_check_gpu.py")], capture_output=True, text=True, env=env,
                       timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "DP_CHECK_OK" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_model_on_second_device_while_first_is_current():
    """dvc.py without --distributed only does `model.to(device)`: a model on cuda:1 driven while cuda:0 is the current
    device must launch on cuda:1 (every entry point runs under the model's device) and give the cuda:0 result."""
    from vidchapters_b200 import Vid2SeqAdam
    fx = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    cfg = fx["cfg"]
    res = []
    for dev in ("cuda:0", "cuda:1"):
        torch.cuda.set_device(0)
        from vidchapters_b200 import Vid2Seq
        tok = Tok(cfg["base_vocab"] + cfg["num_bins"])
        m = Vid2Seq("t5-base", num_features=cfg["num_features"], embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                    heads=cfg["heads"], mlp_dim=cfg["mlp_dim"], vis_drop=0.0, tokenizer=tok, enc_drop=0.0, dec_drop=0.0,
                    num_bins=cfg["num_bins"], t5_config=cfg, seed=0).to(dev)
        m.train()
        opt = Vid2SeqAdam(m, lr=3e-4, clip_max_norm=0.1, world_size=1)
        video, inp, out = fx["video"].to(dev), fx["input_ids"].to(dev), fx["output_ids"].to(dev)
        ld, _ = m(video, {"input_ids": inp, "attention_mask": inp != 0}, {"input_ids": out, "attention_mask": out != 0})
        opt.zero_grad()
        ld["loss"].backward()
        opt.step()
        assert torch.cuda.current_device() == 0
        res.append((ld["loss"].item(), m._params["t5_model.decoder.final_layer_norm.weight"].detach().cpu()))
    assert abs(res[0][0] - res[1][0]) < 1e-5 * abs(res[0][0])
    assert rel(res[1][1], res[0][1]) < 1e-4
