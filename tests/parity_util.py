"""Shared by tests/test_parity_full_gpu.py (CUDA ops, benchmark shapes) and tests/test_host_logic_cpu.py (torch op table,
tiny shapes): teacher-forced per-sub-layer comparison of the engine against the oracle."""
import torch


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def teacher_forced_errors(m, cfg, video, inp, out):
    """Feeds every residual sub-layer of m.engine (and both final norms and the LM head) the ORACLE's own input and
    returns [(rel-L2 of the output, rel-L2 of the residual branch, name)], worst first."""
    from oracle import vid2seq_oracle as O
    eng = m.engine
    m._refresh_shadow()
    ops = eng.ops
    d, H = eng.d, eng.H
    dev = video.device
    B, T = video.shape[0], video.shape[1]
    L, S = inp.shape[1], out.shape[1]
    sd = {k: v.detach() for k, v in m._params.items()}
    trace = []
    with torch.no_grad():
        o = O.vid2seq_forward(sd, cfg, video, inp, inp != 0, out, out != 0, emulate_bf16=True, flash_rounding=True,
                              trace=trace)
    rec = {n: (a, b) for n, a, b in trace}
    E = T + L
    eng._begin_dropout(False)
    bias_e = torch.empty(H, 2 * L - 1, device=dev)
    ops.bias_expand(eng.p(eng.enc_bias_name), eng.lut(L, L, True), bias_e)
    bias_d = torch.empty(H, 2 * S - 1, device=dev)
    ops.bias_expand(eng.p(eng.dec_bias_name), eng.lut(S, S, False), bias_d)
    kmask_e = eng._mask_u8(inp != 0)
    kmask_d = eng._mask_u8(out != 0)
    memory_o, mem_mask_o = rec["memory"]
    memory = memory_o.reshape(B * E, d).to(torch.bfloat16).contiguous()
    mem_mask = mem_mask_o.to(torch.uint8).contiguous()
    errs = []

    def check(tag, x_cuda, x_in, x_out, rows=None):
        xo, xi = x_out.reshape(x_cuda.shape), x_in.reshape(x_cuda.shape)
        if rows is not None:       # rows the caller can observe (see the padded-query note in attn_fwd.cu)
            x_cuda, xo, xi = x_cuda[rows], xo[rows], xi[rows]
        errs.append((rel(x_cuda, xo), rel(x_cuda - xi, xo - xi), tag))

    C = eng.C
    for i, (sa, ff) in enumerate(eng.vit_blocks):
        x_in, x_out = rec[f"vit.{i}.sa"]
        x = eng._sa_fwd(x_in.reshape(B * T, C).contiguous(), sa, B, T, None, None, False, [], dk="vis")
        check(f"vit.{i}.sa", x, x_in, x_out)
        x_in, x_out = rec[f"vit.{i}.ff"]
        x = eng._ff_fwd(x_in.reshape(B * T, C).contiguous(), ff, [], dk="vis")
        check(f"vit.{i}.ff", x, x_in, x_out)
    valid_e = (inp != 0).reshape(-1)
    for i, (sa, ff) in enumerate(eng.enc_blocks):
        p = f"t5_model.encoder.block.{i}."
        x_in, x_out = rec[p + "layer.0.sa"]
        x = eng._sa_fwd(x_in.reshape(B * L, d).contiguous(), sa, B, L, bias_e, kmask_e, False, [], dk="enc")
        check(p + "sa", x, x_in, x_out, rows=valid_e)
        x_in, x_out = rec[p + "layer.1.ff"]
        x = eng._ff_fwd(x_in.reshape(B * L, d).contiguous(), ff, [], dk="enc")
        check(p + "ff", x, x_in, x_out)
    for i, (sa, ca, ff) in enumerate(eng.dec_blocks):
        p = f"t5_model.decoder.block.{i}."
        x_in, x_out = rec[p + "layer.0.sa"]
        x = eng._sa_fwd(x_in.reshape(B * S, d).contiguous(), sa, B, S, bias_d, kmask_d, True, [], dk="dec")
        check(p + "sa", x, x_in, x_out)
        x_in, x_out = rec[p + "layer.1.ca"]
        x = eng._ca_fwd(x_in.reshape(B * S, d).contiguous(), ca, B, S, memory, E, mem_mask, [])
        check(p + "ca", x, x_in, x_out)
        x_in, x_out = rec[p + "layer.2.ff"]
        x = eng._ff_fwd(x_in.reshape(B * S, d).contiguous(), ff, [], dk="dec")
        check(p + "ff", x, x_in, x_out)
    # final norms + head (modeling_t5.py:1113, 1709-1714)
    x_in, x_out = rec["enc.final"]
    y = torch.empty(B * L, d, device=dev, dtype=torch.bfloat16)
    ops.norm_fwd(0, x_in.reshape(B * L, d).contiguous(), eng.pv("t5_model.encoder.final_layer_norm.weight"), None,
                 out_bf16=y, eps=1e-6)
    errs.append((rel(y, x_out.reshape(B * L, d).to(torch.bfloat16)), 0.0, "enc.final_norm(bf16)"))
    x_in, x_out = rec["dec.final"]
    seq = torch.empty(B * S, d, device=dev, dtype=torch.bfloat16)
    ops.norm_fwd(0, x_in.reshape(B * S, d).contiguous(), eng.pv("t5_model.decoder.final_layer_norm.weight"), None,
                 out_bf16=seq, eps=1e-6, out_scale=d ** -0.5)
    Vp = (eng.V + 7) // 8 * 8
    logits = torch.empty(B * S, Vp, device=dev)[:, :eng.V]
    ops.gemm(seq, eng.pb("t5_model.shared.weight"), logits)
    errs.append((rel(logits, o["logits"].reshape(B * S, -1)), 0.0, "final_norm+lm_head"))
    errs.sort(reverse=True)
    return errs


# Degenerate batches shared by the oracle-vs-reference and the engine-vs-oracle edge-case tests
EDGE_CASES = {
    # B, T, L, S, mutate(inp, out)
    "batch1_single_target_token": (1, 10, 9, 1, lambda inp, out: None),
    "one_frame_one_text_token": (2, 1, 1, 5, lambda inp, out: None),
    "empty_asr_row": (2, 10, 12, 6, lambda inp, out: (inp.__setitem__((0, slice(1, None)), 0), inp.__setitem__((0, 0), 1))),   # "</s>" only
    "target_row_of_eos_only": (2, 10, 12, 6, lambda inp, out: (out.__setitem__((1, slice(1, None)), 0), out.__setitem__((1, 0), 1))),
    "ragged_everything": (3, 7, 33, 19, lambda inp, out: (inp.__setitem__((0, slice(5, None)), 0), inp.__setitem__((2, slice(30, None)), 0),
                                                        out.__setitem__((0, slice(2, None)), 0), out.__setitem__((1, slice(18, None)), 0))),
}


def edge_batch(name, seed=21):
    B, T, L, S, mut = EDGE_CASES[name]
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(B, T, 768, generator=g)
    inp = torch.randint(2, 1100, (B, L), generator=g)
    out = torch.randint(2, 1100, (B, S), generator=g)
    mut(inp, out)
    return video, inp, out
