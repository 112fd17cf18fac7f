"""End-to-end parity of the CUDA path (through the drop-in module and the C ABI) on the GPU.

Tier A: vs the bf16-operand emulation oracle on the same device evaluated at the kernels' rounding points
        (flash_rounding=True: un-normalised probabilities rounded to bf16).  Gate = max(1e-3, 1.5 x the MEASURED floor),
        the floor being the distance between two legal evaluations of that arithmetic (fp32- vs fp64-accumulated);
        time-token argmax bit-exact except ties inside the oracle's own noise band.  tests/test_parity_full_gpu.py
        repeats this at the benchmark shapes and proves <= 1e-3 per sub-layer (teacher-forced).
Tier B: vs the golden vectors minted from the real fp32 reference: logits rel-L2 must stay below the reference's own
        bf16 error (1.2e-2), loss within 2e-3 rel, argmax mismatches only at ties of the reference itself.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


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


@pytest.mark.parametrize("name", ["tiny", "tiny_proj", "tiny_long", "t5base_cfg1"])
def test_cuda_vs_golden_and_emulation_oracle(name):
    from oracle import vid2seq_oracle as O
    fx = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    cfg = fx["cfg"]
    m = build(cfg)
    video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
    it = {"input_ids": inp, "attention_mask": inp != 0}
    ot = {"input_ids": out, "attention_mask": out != 0}
    loss, logits = m.forward_logits(video, it, ot)
    V0 = cfg["base_vocab"]
    # ---- tier B: the real reference's fp32 outputs
    assert abs(loss.item() - fx["loss"].item()) < 2e-3 * abs(fx["loss"].item())
    if "logits" in fx:
        e = rel(logits.cpu(), fx["logits"])
    else:
        e = rel(logits[..., V0:].cpu(), fx["logits_time"])
    print(f"[{name}] tier-B logits rel-L2 vs fp32 reference: {e:.3e}")
    assert e < 1.2e-2
    agree = (logits[..., V0:].argmax(-1).cpu() == fx["time_argmax"]).float().mean().item()
    print(f"[{name}] time-token argmax agreement vs fp32 reference: {agree:.4f} (every mismatch must be a tie, below)")
    # ---- tier A: emulation oracle on the same device, gated by the MEASURED noise floor of that arithmetic:
    # emu = bf16 operands / fp32 accumulate at the kernels' rounding points, emu64 = the same bf16 operands with every
    # product accumulated in fp64 (accumulation-order free).  |emu - emu64| is what two legal evaluations of the same
    # arithmetic differ by (a 1-ulp fp32 difference flips bf16 roundings downstream; the flips amplify through the
    # layers — the teacher-forced per-sub-layer test in test_parity_full_gpu.py shows each kernel alone is ~1e-4).
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m._params.items()}
    o = O.vid2seq_forward(sd, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True)
    sd64 = {k: v.detach().clone().requires_grad_(True) for k, v in m._params.items()}
    o64 = O.vid2seq_forward(sd64, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True,
                            acc64=True)
    z, z64 = o["logits"].detach(), o64["logits"].detach()
    floor = rel(z, z64)
    ea = rel(logits, z64)
    print(f"[{name}] tier-A logits rel-L2 vs emu64: {ea:.3e}; measured floor |emu - emu64|: {floor:.3e}; vs emu: {rel(logits, z):.3e}")
    assert ea <= max(1e-3, 1.5 * floor), (ea, floor)
    # time-token argmax: exact, except ties inside the oracle's own noise band
    tz, tz64, tzc = z[..., V0:], z64[..., V0:], logits[..., V0:]
    noise = (tz - tz64).abs().max().item()
    top2 = tz64.topk(2, dim=-1).values
    mism = tzc.argmax(-1) != tz64.argmax(-1)
    if bool(mism.any()):
        gap = (top2[..., 0] - top2[..., 1])[mism].max().item()
        print(f"[{name}] {int(mism.sum())} time-token argmax ties: max oracle gap {gap:.3e}, oracle noise {noise:.3e}")
        assert gap <= 3.0 * noise, (gap, noise)
    # the same tie rule against the real fp32 reference: a mismatch is legal only where the reference's own top-2 gap is
    # inside the band by which the bf16-operand arithmetic (emu64) moves the reference's logits
    gt = (fx["logits"][..., V0:] if "logits" in fx else fx["logits_time"]).to(tzc.device)
    noise_b = (tz64 - gt).abs().max().item()
    mism_b = tzc.argmax(-1) != gt.argmax(-1)
    if bool(mism_b.any()):
        t2 = gt.topk(2, dim=-1).values
        gap_b = (t2[..., 0] - t2[..., 1])[mism_b].max().item()
        print(f"[{name}] {int(mism_b.sum())} argmax ties vs fp32: max reference gap {gap_b:.3e}, bf16-arithmetic band {noise_b:.3e}")
        assert gap_b <= 3.0 * noise_b, (gap_b, noise_b)
    assert abs(loss.item() - o64["loss"].item()) <= max(5e-4, 3 * abs(o["loss"].item() - o64["loss"].item()) /
                                                         abs(o64["loss"].item())) * abs(o64["loss"].item())
    # ---- backward through the module surface
    m.train()
    ld, vd = m(video, it, ot)
    ld["loss"].backward()
    o["loss"].backward()
    o64["loss"].backward()
    errs, ferrs, nerrs = [], [], []
    for n, p in m._params.items():
        errs.append((rel(p.grad, sd64[n].grad), n))
        ferrs.append((rel(sd[n].grad, sd64[n].grad), n))
        gn = fx["grad_norms"][n]
        nerrs.append((abs(p.grad.norm().item() - gn) / (gn + 1e-12), n))
    errs.sort(reverse=True)
    ferrs.sort(reverse=True)
    nerrs.sort(reverse=True)
    med, fmed = errs[len(errs) // 2][0], ferrs[len(ferrs) // 2][0]
    print(f"[{name}] per-parameter gradient rel-L2 vs emu64: worst {errs[0][0]:.3e} ({errs[0][1]}), median {med:.3e}; "
          f"floor |emu - emu64|: worst {ferrs[0][0]:.3e}, median {fmed:.3e}; "
          f"gradient-norm error vs the real reference: worst {nerrs[0][0]:.3e} ({nerrs[0][1]})")
    # gradients: the floor is ONE draw of the noise (emu vs emu64) and so is our error; on the 2-layer models the ratio of
    # two draws reaches 1.6 (measured, profiles/r02_parity.json), so 2x here; the benchmark-shape tests hold 1.5x
    assert med <= max(1e-3, 2.0 * fmed), (med, fmed)
    assert errs[0][0] <= max(1e-3, 2.0 * ferrs[0][0]), (errs[:3], ferrs[:3])
    assert nerrs[0][0] < 6e-2, nerrs[:3]


def test_train_steps_reduce_loss_and_match_oracle_tail():
    from oracle import vid2seq_oracle as O
    from vidchapters_b200 import TINY, Vid2SeqAdam
    fx = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    cfg = fx["cfg"]
    m = build(cfg)
    opt = Vid2SeqAdam(m, lr=3e-4, clip_max_norm=0.1, world_size=1)
    video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
    it = {"input_ids": inp, "attention_mask": inp != 0}
    ot = {"input_ids": out, "attention_mask": out != 0}
    before = {n: p.detach().clone() for n, p in m._params.items()}
    losses = []
    for step in range(4):
        ld, _ = m(video, it, ot)
        opt.zero_grad()
        ld["loss"].backward()
        if step == 0:
            grads = {n: p.grad.clone() for n, p in m._params.items()}
        opt.step()
        if step == 0:
            params = {k: v.clone() for k, v in before.items()}
            O.clip_adam_renorm_(params, grads, {}, lr=3e-4, clip_max_norm=0.1, num_bins=cfg["num_bins"])
            for n, p in m._params.items():
                assert rel(p.detach() - before[n], params[n] - before[n]) < 2e-3, n
            # the real reference's own post-step tensors
            assert rel(m._params["t5_model.shared.weight"][-100:].cpu(), fx["after_step"]["time_rows"]) < 1e-3
        losses.append(ld["loss"].item())
    assert losses[-1] < losses[0], losses


def test_two_pass_and_stock_optimizer():
    """dvc.py default step: generative + denoising pass sharing video_dict; stock torch Adam + clip_grad_norm_."""
    from vidchapters_b200 import TINY
    cfg = dict(TINY, num_features=10)
    m = build(cfg)
    g = torch.Generator().manual_seed(5)
    B, T = 2, 10
    video = torch.randn(B, T, 768, generator=g).cuda()
    mk = lambda L: torch.randint(2, 1100, (B, L), generator=g).cuda()
    i1, o1, i2, o2 = mk(40), mk(20), mk(30), mk(16)
    opt = torch.optim.Adam(m.parameters(), lr=3e-4)
    l1, vd = m(video, {"input_ids": i1, "attention_mask": i1 != 0}, {"input_ids": o1, "attention_mask": o1 != 0})
    l2, _ = m(vd, {"input_ids": i2, "attention_mask": i2 != 0}, {"input_ids": o2, "attention_mask": o2 != 0})
    opt.zero_grad()
    (l1["loss"] + l2["loss"]).backward()
    gn = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    assert torch.isfinite(gn)
    opt.step()
    l3, _ = m(video, {"input_ids": i1, "attention_mask": i1 != 0}, {"input_ids": o1, "attention_mask": o1 != 0})
    assert l3["loss"].item() < l1["loss"].item()


def test_dropout_training_step_cuda_vs_torch_ops():
    """Reference default dropout 0.1 in training mode: the CUDA engine and the torch op table regenerate the same
    counter-based masks, so loss and gradients must agree to bf16 noise."""
    from oracle.torch_ops import TorchOps
    from vidchapters_b200.engine import Vid2SeqEngine
    from vidchapters_b200.ops import CudaOps
    fx = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    cfg = fx["cfg"]
    m = build(cfg)
    eng_c = m.engine
    eng_t = Vid2SeqEngine(cfg, TorchOps(), "cuda")
    eng_t.flat_p.copy_(eng_c.flat_p)
    eng_c.sync_bf16(); eng_t.sync_bf16()
    video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
    res = []
    for eng in (eng_c, eng_t):
        eng.drop_rates = dict(vis=0.1, enc=0.1, dec=0.1)
        eng.drop_seed, eng._drop_calls = 5, 0
        loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0, training=True)
        eng.zero_grad()
        eng.backward(ctx)
        res.append((loss.item(), eng.flat_g.clone()))
    assert abs(res[0][0] - res[1][0]) < 2e-3 * abs(res[1][0]), (res[0][0], res[1][0])
    worst = 0.0
    for n in eng_c.layout:
        o, shp, k = eng_c.layout[n]
        e = rel(res[0][1][o:o + k], res[1][1][o:o + k])
        worst = max(worst, e)
        assert e < 8e-2, (n, e)
    print(f"[dropout] loss cuda {res[0][0]:.5f} torch-ops {res[1][0]:.5f}; worst gradient rel-L2 {worst:.3e}")
    # module surface: dropout active in train(), off in eval()
    m.vis_drop = m.enc_drop = m.dec_drop = 0.1
    it = {"input_ids": inp, "attention_mask": inp != 0}
    ot = {"input_ids": out, "attention_mask": out != 0}
    m.train(); l_tr = m(video, it, ot)[0]["loss"].item()
    m.eval(); l_ev = m(video, it, ot)[0]["loss"].item()
    assert abs(l_tr - l_ev) > 1e-3 and abs(l_ev - fx["loss"].item()) < 2e-3 * abs(l_ev)


def test_greedy_generate_cuda_matches_oracle():
    """Vid2Seq.generate (num_beams=1): KV-cache decode with the CUDA-graphed step vs the uncached oracle greedy loop."""
    from oracle import vid2seq_oracle as O
    fx = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    cfg = fx["cfg"]
    m = build(cfg)
    m.eval()

    class TokD(Tok):
        def batch_decode(self, ids, skip_special_tokens=True):
            return [" ".join(str(int(t)) for t in row if not (skip_special_tokens and int(t) in (0, 1))) for row in ids]

    m.t5_tokenizer = TokD(cfg["base_vocab"] + cfg["num_bins"])
    video, inp = fx["video"].cuda(), fx["input_ids"].cuda()
    texts = m.generate(video, {"input_ids": inp, "attention_mask": inp != 0}, num_beams=1, max_length=12)
    seq = m.last_generated_ids
    assert len(texts) == video.shape[0] and seq.shape[1] <= 13 and bool((seq[:, 0] == 0).all())
    sd = {k: v.detach() for k, v in m._params.items()}
    memory, mem_mask, B, E = m.engine.encode(video, inp, inp != 0)
    ref = O.greedy_decode(sd, cfg, memory.float().view(B, E, -1), mem_mask.long(), max_new_tokens=12, emulate_bf16=True)
    n = min(seq.shape[1], ref.shape[1])
    agree = (seq[:, :n] == ref[:, :n]).float().mean().item()
    print(f"[generate] greedy ids agreement with the oracle: {agree:.3f}  ids[0]={seq[0].tolist()}")
    assert agree == 1.0
    with pytest.raises(NotImplementedError):
        m.generate(video, {"input_ids": inp, "attention_mask": inp != 0}, num_beams=4, use_nucleus_sampling=True)


def test_beam_search_generate_cuda_matches_oracle():
    """Vid2Seq.generate with the reference's default num_beams=4 (and greedy) on a model that has learnt something:
    KV-cache decode, vc_beam_topk / vc_kv_reorder and the host n-best bookkeeping vs the uncached oracle restatement of
    HF-4.28 beam search (oracle/vid2seq_oracle.py::beam_search_decode; parity of THAT against HF is unpinned)."""
    from oracle import vid2seq_oracle as O
    from vidchapters_b200 import Vid2SeqAdam
    fx = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    cfg = fx["cfg"]
    video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
    it = {"input_ids": inp, "attention_mask": inp != 0}
    ot = {"input_ids": out, "attention_mask": out != 0}

    class TokD(Tok):
        def batch_decode(self, ids, skip_special_tokens=True):
            return [" ".join(str(int(t)) for t in row if not (skip_special_tokens and int(t) in (0, 1))) for row in ids]

    for steps in (10, 40):          # a half-trained (flat distributions) and a memorised model
        m = build(cfg)
        m.train()
        opt = Vid2SeqAdam(m, lr=2e-3, clip_max_norm=1.0, world_size=1)
        for _ in range(steps):
            ld, _ = m(video, it, ot); opt.zero_grad(); ld["loss"].backward(); opt.step()
        m.eval()
        m.t5_tokenizer = TokD(cfg["base_vocab"] + cfg["num_bins"])
        sd = {k: v.detach().clone() for k, v in m._params.items()}
        memory, mem_mask, B, E = m.engine.encode(video, inp, inp != 0)
        mem32 = memory.float().view(B, E, -1)
        for nb in (1, 4):
            m.generate(video, it, num_beams=nb, max_length=16)
            seq = m.last_generated_ids
            ref = (O.greedy_decode if nb == 1 else lambda *a, **k: O.beam_search_decode(*a, num_beams=nb, **k))(
                sd, cfg, mem32, mem_mask.long(), max_new_tokens=16, emulate_bf16=True).to(seq.device)
            n = min(seq.shape[1], ref.shape[1])
            agree = (seq[:, :n] == ref[:, :n]).float().mean().item()
            print(f"[generate] steps={steps} num_beams={nb}: ids agreement with the oracle {agree:.3f}; ids[0]={seq[0].tolist()}")
            # flat distributions (10 steps) may break a near-tie differently under bf16 rounding flips; the memorised model may not
            assert agree >= (1.0 if steps == 40 else 0.7) and bool((seq[:, n:] == 0).all()) and bool((ref[:, n:] == 0).all())
        if steps == 40:             # the memorised model reproduces its targets
            tgt = out[0][out[0] != 0]
            assert seq[0, 1:1 + len(tgt)].tolist() == tgt.tolist()


def test_graphed_train_step_matches_eager_and_redraws_dropout():
    from vidchapters_b200 import GraphedTrainStep, Vid2SeqAdam
    fx = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    cfg = fx["cfg"]
    video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
    it = {"input_ids": inp, "attention_mask": inp != 0}
    ot = {"input_ids": out, "attention_mask": out != 0}
    # eager reference run (no dropout): 4 steps
    m1 = build(cfg); m1.train()
    o1 = Vid2SeqAdam(m1, lr=3e-4, clip_max_norm=0.1, world_size=1)
    eager = []
    for _ in range(4):
        ld, _ = m1(video, it, ot); o1.zero_grad(); ld["loss"].backward(); o1.step(); eager.append(ld["loss"].item())
    # graphed run from the same init: warm-up step eager (1), then captured replays (3)
    m2 = build(cfg); m2.train()
    o2 = Vid2SeqAdam(m2, lr=3e-4, clip_max_norm=0.1, world_size=1)
    ld, _ = m2(video, it, ot); o2.zero_grad(); ld["loss"].backward(); o2.step()
    g = GraphedTrainStep(m2, o2, video, inp, out, warmup_steps=0)
    graphed = [ld["loss"].item()] + [g(video, inp, out).item() for _ in range(3)]
    # same arithmetic, different accumulation order of the fp32 atomics: the first steps agree to ~1e-5; by the 4th Adam
    # step of this tiny model (loss 15 -> 1.9) the difference has been amplified to several 1e-3
    for i, (a, b) in enumerate(zip(eager, graphed)):
        assert abs(a - b) < (5e-3 if i < 3 else 3e-2) * abs(a), (eager, graphed)
    # with dropout the same batch gives a different loss on every replay (device-side salt), still finite/decreasing
    m3 = build(cfg); m3.train()
    m3.vis_drop = m3.enc_drop = m3.dec_drop = 0.1
    o3 = Vid2SeqAdam(m3, lr=3e-4, clip_max_norm=0.1, world_size=1)
    ld, _ = m3(video, it, ot); o3.zero_grad(); ld["loss"].backward(); o3.step()
    g3 = GraphedTrainStep(m3, o3, video, inp, out, warmup_steps=0)
    ls = [g3().item() for _ in range(4)]
    assert len(set(round(x, 4) for x in ls)) == 4 and all(x == x for x in ls), ls
    g.close()
    g3.close()


def test_full_size_properties_config2():
    """BASELINE configs[1] at full size (t5-base, batch 16, 100 frames, 1000 ASR tokens, 256 target tokens): too big
    for the CPU oracle, so parity is checked through properties the reference's arithmetic guarantees.
      P1 right-padding invariance: pad keys carry the additive finfo.min mask (modeling_t5.py:539-580) and T5's bias is
         relative, so truncating trailing padding (1000 -> 896 tokens: one key tile less, different tile skipping) must
         not change the loss;
      P2 batch-permutation equivariance of the mean loss and of the gradient;
      P3 linearity of the backward in the upstream gradient (loss * 2 -> gradients * 2);
      P4 run-to-run determinism of the forward."""
    from vidchapters_b200 import T5_BASE
    cfg = dict(T5_BASE)
    m = build(cfg)
    m.train()
    g = torch.Generator().manual_seed(77)
    B, T, L, S = 16, 100, 1000, 256
    video = torch.randn(B, T, 768, generator=g).cuda()
    inp = torch.zeros(B, L, dtype=torch.long)
    out = torch.zeros(B, S, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(500, 871, (1,), generator=g))          # all rows fit in 896 tokens
        inp[b, :n] = torch.randint(2, 32100, (n,), generator=g); inp[b, n - 1] = 1
        k = int(torch.randint(128, 257, (1,), generator=g))
        out[b, :k] = torch.randint(2, 32200, (k,), generator=g); out[b, k - 1] = 1
    inp, out = inp.cuda(), out.cuda()
    tok = lambda x: {"input_ids": x, "attention_mask": x != 0}

    def run(v, i, o, scale=1.0):
        for p in m.parameters():
            p.grad = None
        ld, _ = m(v, tok(i), tok(o))
        (ld["loss"] * scale).backward()
        return ld["loss"].item(), m.engine.flat_g.clone()

    l0, g0 = run(video, inp, out)
    assert l0 == l0 and 1.0 < l0 < 40.0
    # P4: the forward is deterministic; the backward accumulates dQ / weight gradients with fp32 atomics (TMA reduce-add)
    # in arrival order and the bf16 roundings downstream of them flip, so gradients repeat to ~3e-3, not bit for bit
    l0b, g0b = run(video, inp, out)
    floor = rel(g0b, g0)
    print(f"[config2 full size] run-to-run: loss {abs(l0b - l0) / l0:.2e}, gradient {floor:.3e}")
    assert abs(l0b - l0) < 2e-6 * l0 and floor < 1e-2
    # P1
    l1, g1 = run(video, inp[:, :896].contiguous(), out)
    print(f"[config2 full size] loss {l0:.6f}; truncated padding {l1:.6f}; grad rel {rel(g1, g0):.3e}")
    assert abs(l1 - l0) < 2e-3 * l0
    assert rel(g1, g0) < 6e-2          # bf16 rounding-flip floor of a 12+12-layer model (see the tier-A notes above)
    # P2
    perm = torch.randperm(B, generator=g).cuda()
    l2, g2 = run(video[perm], inp[perm], out[perm])
    assert abs(l2 - l0) < 1e-5 * l0 and rel(g2, g0) < max(1e-2, 3 * floor), (l0, l2, rel(g2, g0))
    # P3
    l3, g3 = run(video, inp, out, scale=2.0)
    assert rel(g3, 2.0 * g0) < max(1e-2, 3 * floor)


@pytest.mark.parametrize("use_video,use_speech", [(True, False), (False, True)], ids=["no_speech", "no_video"])
def test_modality_variants_cuda(use_video, use_speech):
    """--no_speech / --no_video (vid2seq.py:59-84) through the drop-in module on the GPU vs the oracle."""
    from oracle import vid2seq_oracle as O
    from vidchapters_b200 import Vid2Seq
    fx = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    cfg = fx["cfg"]
    tok = Tok(cfg["base_vocab"] + cfg["num_bins"])
    m = Vid2Seq("t5-base", num_features=cfg["num_features"], embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                heads=cfg["heads"], mlp_dim=cfg["mlp_dim"], vis_drop=0.0, tokenizer=tok, enc_drop=0.0, dec_drop=0.0,
                num_bins=cfg["num_bins"], t5_config=cfg, seed=0, use_video=use_video, use_speech=use_speech).to("cuda")
    m.train()
    video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
    ld, vd = m(video, {"input_ids": inp, "attention_mask": inp != 0}, {"input_ids": out, "attention_mask": out != 0})
    ld["loss"].backward()
    assert (vd is None) == (not use_video)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m._params.items()}
    o = O.vid2seq_forward(sd, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True,
                          use_video=use_video, use_speech=use_speech)
    o["loss"].backward()
    assert abs(ld["loss"].item() - o["loss"].item()) < 1e-3 * abs(o["loss"].item())
    errs = []
    for n, p in m._params.items():
        if sd[n].grad is None:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, n
        else:
            errs.append((rel(p.grad, sd[n].grad), n))
    errs.sort(reverse=True)
    print(f"[variant video={use_video} speech={use_speech}] loss {ld['loss'].item():.5f} vs oracle {o['loss'].item():.5f}; worst grad rel {errs[0]}")
    assert errs[0][0] < 1.2e-1 and errs[len(errs) // 2][0] < 4e-2


@pytest.mark.parametrize("flag,value", [("VIDCHAP_FUSE_CROSS_KV", "0")])
def test_engine_switches_match_default_path(flag, value, monkeypatch):
    """The engine switch that changes launch structure and the parameter layout (default: the cross-attention K/V
    projections of ALL decoder layers as one GEMM over the grouped layout; =0: one GEMM pair per layer): same forward
    bit for bit, gradients equal up to fp32 summation order — eagerly AND when the step is replayed as a CUDA graph."""
    from vidchapters_b200 import GraphedTrainStep, Vid2SeqAdam
    fx = torch.load(os.path.join(GOLD, "tiny_long.pt"), weights_only=False)
    cfg = fx["cfg"]
    video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
    it = {"input_ids": inp, "attention_mask": inp != 0}
    ot = {"input_ids": out, "attention_mask": out != 0}

    def run():
        m = build(cfg)
        m.train()
        opt = Vid2SeqAdam(m, lr=3e-4, clip_max_norm=0.1, world_size=1)
        ld, _ = m(video, it, ot)
        opt.zero_grad()
        ld["loss"].backward()
        grads = {n: p.grad.detach().clone() for n, p in m._params.items()}
        opt.step()
        g = GraphedTrainStep(m, opt, video, inp, out, warmup_steps=0)
        losses = [ld["loss"].item()] + [g(video, inp, out).item() for _ in range(3)]
        torch.cuda.synchronize()
        g.close()
        return losses, grads, {n: p.detach().clone() for n, p in m._params.items()}

    l0, g0, p0 = run()
    l0b, g0b, p0b = run()                      # run-to-run noise of the default path (atomics)
    monkeypatch.setenv(flag, value)
    l1, g1, p1 = run()
    assert abs(l1[0] - l0[0]) <= 1e-6 * abs(l0[0]), (l0, l1)
    worst, noise = 0.0, 0.0
    for n in g0:
        worst = max(worst, rel(g1[n], g0[n]))
        noise = max(noise, rel(g0b[n], g0[n]))
    print(f"[{flag}] losses default {l0} vs switched {l1}; worst per-parameter gradient rel-L2 {worst:.3e} "
          f"(run-to-run noise of the default path {noise:.3e})")
    # (one K = layers*2*inner GEMM vs 12 accumulating ones: a different fp32 summation order of d(memory), then the
    #  bf16 rounding flips of everything downstream — the same mechanism as the tier-A floor, here ~4e-3 on this model)
    assert worst <= 1.5e-2, (worst, noise)
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 2e-2 * abs(a), (l0, l1)
    perr = max(rel(p1[n] - fx_p, p0[n] - fx_p) if (p0[n] - fx_p).norm() > 0 else 0.0
               for n, fx_p in ((n, build_init(cfg)[n]) for n in ("t5_model.decoder.block.0.layer.1.EncDecAttention.k.weight",
                                                                 "visual_encoder.blocks.0.attn.qkv.weight")))
    assert perr < 0.2, perr                   # 4 Adam steps from the same init land on the same updates


def build_init(cfg, _cache={}):
    from vidchapters_b200.init import init_state_dict
    key = id(cfg)
    if key not in _cache:
        _cache[key] = {k: v.cuda() for k, v in init_state_dict(cfg, 0).items()}
    return _cache[key]


def test_generate_options_cuda_match_oracle():
    """Vid2Seq.generate's remaining kwargs on the CUDA path (vid2seq.py:100-167 -> HF generate): repetition_penalty and
    min_length (greedy and beam search), num_captions (n best beams), nucleus sampling (top_p -> 0 must reproduce greedy;
    a fixed torch seed reproduces itself; every sampled row is a valid sequence)."""
    from oracle import vid2seq_oracle as O
    from vidchapters_b200 import Vid2SeqAdam
    fx = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    cfg = fx["cfg"]
    video, inp, out = fx["video"].cuda(), fx["input_ids"].cuda(), fx["output_ids"].cuda()
    it = {"input_ids": inp, "attention_mask": inp != 0}
    ot = {"input_ids": out, "attention_mask": out != 0}

    class TokD(Tok):
        def batch_decode(self, ids, skip_special_tokens=True):
            return [" ".join(str(int(t)) for t in row if not (skip_special_tokens and int(t) in (0, 1))) for row in ids]

    m = build(cfg)
    m.train()
    opt = Vid2SeqAdam(m, lr=2e-3, clip_max_norm=1.0, world_size=1)
    for _ in range(40):
        ld, _ = m(video, it, ot); opt.zero_grad(); ld["loss"].backward(); opt.step()
    m.eval()
    m.t5_tokenizer = TokD(cfg["base_vocab"] + cfg["num_bins"])
    sd = {k: v.detach().clone() for k, v in m._params.items()}
    memory, mem_mask, B, E = m.engine.encode(video, inp, inp != 0)
    mem32 = memory.float().view(B, E, -1)
    for kw in (dict(repetition_penalty=2.0), dict(min_length=14), dict(repetition_penalty=1.3, min_length=10)):
        m.generate(video, it, num_beams=1, max_length=16, **kw)
        seq = m.last_generated_ids
        ref = O.greedy_decode(sd, cfg, mem32, mem_mask.long(), max_new_tokens=16, emulate_bf16=True, **kw).to(seq.device)
        n = min(seq.shape[1], ref.shape[1])
        assert torch.equal(seq[:, :n], ref[:, :n]), (kw, seq, ref)
        m.generate(video, it, num_beams=4, max_length=16, **kw)
        seq = m.last_generated_ids
        ref = O.beam_search_decode(sd, cfg, mem32, mem_mask.long(), num_beams=4, max_new_tokens=16, emulate_bf16=True, **kw).to(seq.device)
        n = min(seq.shape[1], ref.shape[1])
        assert torch.equal(seq[:, :n], ref[:, :n]), (kw, seq, ref)
    texts = m.generate(video, it, num_beams=4, max_length=16, num_captions=3)
    assert len(texts) == 3 * B
    ref = O.beam_search_decode(sd, cfg, mem32, mem_mask.long(), num_beams=4, max_new_tokens=16, emulate_bf16=True, num_return=3)
    seq = m.last_generated_ids
    n = min(seq.shape[1], ref.shape[1])
    assert torch.equal(seq[:, :n].cpu(), ref[:, :n])
    m.generate(video, it, num_beams=1, max_length=16)
    greedy = m.last_generated_ids.clone()
    m.generate(video, it, use_nucleus_sampling=True, num_beams=0, top_p=1e-4, max_length=16)     # nucleus = the argmax only
    n = min(greedy.shape[1], m.last_generated_ids.shape[1])
    assert torch.equal(m.last_generated_ids[:, :n], greedy[:, :n])
    torch.manual_seed(5)
    t1 = m.generate(video, it, use_nucleus_sampling=True, num_beams=0, top_p=0.95, temperature=2.0, max_length=16, num_captions=2)
    s1 = m.last_generated_ids.clone()
    torch.manual_seed(5)
    m.generate(video, it, use_nucleus_sampling=True, num_beams=0, top_p=0.95, temperature=2.0, max_length=16, num_captions=2)
    assert len(t1) == 2 * B and torch.equal(s1, m.last_generated_ids)
    assert bool((s1[:, 0] == 0).all()) and int(s1.max()) < cfg["base_vocab"] + cfg["num_bins"]
