"""bench.py — Vid2Seq on B200 (BASELINE.json metric: train-step tokens/sec, t5-base; secondary rows: t5-large, greedy decode).

    python bench.py --gpus 1 --steps K --warmup W             # default: BASELINE configs[1] (t5-base, batch 16) train step
    python bench.py --model t5-large [--gpus N]               # configs[3] (batch 8 per GPU)
    python bench.py --mode decode                             # configs[4]: greedy decode, batch 64, 256 new tokens
    python bench.py --impl reference [...]                    # the reference algorithm's CPU path on this box's host cores
  (torchrun for N > 1: one process per GPU; RANK / LOCAL_RANK / WORLD_SIZE from the environment.)

Train: one "step" = one pass of dvc.py:42-133 over one synthetic batch: H2D of the batch, Vid2Seq forward (generative
pass), backward, [gradient all-reduce + the loss-scalar all-reduce of dvc.py:103], clip + Adam + time-token renorm, loss
read-back.  tokens/step/GPU = B * (T + L + S).
Decode: one "step" = one `generate` call (visual + text encoder, cross K/V projection, 256 greedy steps); tokens = B * 256.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, device-timed (CUDA events, max over ranks).  `e2e`: the
same through the public API with pinned HOST buffers copied inside the timed region and the result read back every step.
`roofline`: train = all tcgen05 GEMM launches of one instrumented step (2*M*N*K flops / CUDA-event time) against
MEASURED_PEAKS.json's sustained bf16 figure; decode = algorithmic bytes per decode step / its time against hbm_gbs.
`cpu_baseline` and `--impl reference`: the oracle port of the reference (oracle/vid2seq_oracle.py, pinned to the real
reference by tests/) on the host cores, same shapes, dropout as the GPU arm, a bounded number of videos per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_FRAMES, L_ASR, S_TGT = 100, 1000, 256
MODEL_BATCH = {"t5-base": 16, "t5-large": 8}     # per-GPU batch BASELINE.json quotes (configs[1] / configs[3])
DECODE_BATCH, DECODE_NEW = 64, 256               # configs[4]
# algorithmic forward GFLOP per sample (2*MAC, full S x S for causal); a train step = 3x
FWD_GFLOP_PER_SAMPLE = {"t5-base": 327.65, "t5-large": 8465.0 / 8}
REF_VIDEOS = 4                                   # videos per step of the CPU arms (bounded sample of the GPU batch)


class Tok:
    pad_token_id, eos_token_id = 0, 1

    def __init__(self, n=32200):
        self.n = n

    def __len__(self):
        return self.n

    def batch_decode(self, ids, skip_special_tokens=True):
        return ["" for _ in ids]


def synth_batch(B, T, L, S, seed, base_vocab=32100, vocab=32200):
    """SURVEY §8d synthetic inputs: video ~ N(0,1); ASR ids U{2..V-1}, per-row valid length U{L/2..L}, eos, 0-pad;
    targets [time,time,text*k]... with time ids in [32100,32200), eos, 0-pad, valid length U{S/2..S}."""
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(B, T, 768, generator=g)
    inp = torch.randint(2, vocab, (B, L), generator=g)
    out = torch.randint(2, base_vocab, (B, S), generator=g)
    tpos = torch.arange(S) % 8 < 2
    out[:, tpos] = torch.randint(base_vocab, vocab, (B, int(tpos.sum())), generator=g)
    for b in range(B):
        li = int(torch.randint(L // 2, L + 1, (1,), generator=g))
        lo = int(torch.randint(S // 2, S + 1, (1,), generator=g))
        inp[b, li - 1] = 1
        inp[b, li:] = 0
        out[b, lo - 1] = 1
        out[b, lo:] = 0
    return video, inp, out


def model_cfg(name):
    from vidchapters_b200.config import T5_BASE, T5_LARGE
    return dict(T5_BASE if name == "t5-base" else T5_LARGE)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx, self.rows, self.proc, self.t_mark = gpu_index, [], None, 0.0

    def mark(self):
        """Start of the timed region: only samples read after this count.  (The sampler itself is started BEFORE the
        warm-up: nvidia-smi's start-up enumerates the devices under the driver lock and stalled the first timed
        generate() call — which allocates, captures and instantiates a graph — by up to seconds when it overlapped.)"""
        self.t_mark = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.rows = [r for t, r in self.rows if t >= self.t_mark]
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"tf_sustained": d.get("bf16_tflops_sustained", 1400.0), "tf_burst": d.get("bf16_tflops", 1590.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tf_sustained": 1400.0, "tf_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def committed_profile(name):
    """Numbers that only ncu can give (DRAM traffic of the dominant kernel, per-kernel tensor-pipe %) are read from the
    committed summary of the capture they came from (profiles/<name>), never measured under the profiler here."""
    p = os.path.join(ROOT, "profiles", name)
    if os.path.isfile(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------- CPU arms (oracle port of the reference)
def _pick_threads():
    """The reference step is hundreds of medium-sized fp32 ops: on a many-core host all cores are not always fastest
    (OpenMP barriers).  VIDCHAP_CPU_THREADS pins it; otherwise the arm calibrates {all cores, 32} on one step each."""
    env = os.environ.get("VIDCHAP_CPU_THREADS")
    if env:
        return [int(env)]
    n = os.cpu_count() or 1
    return sorted({n, min(n, 32)}, reverse=True)


def cpu_train_step_time(cfg, B, T, L, S, steps, warmup, budget_s, dropout):
    from oracle import vid2seq_oracle as O
    from vidchapters_b200.init import init_state_dict
    sd = init_state_dict(cfg, 0)
    params = {k: v.requires_grad_(True) for k, v in sd.items()}
    state = {}
    video, inp, out = synth_batch(B, T, L, S, 1, cfg["base_vocab"], cfg["base_vocab"] + cfg["num_bins"])
    tdrop = dict(vis=dropout, enc=dropout, dec=dropout) if dropout > 0 else None

    def one_step():
        t0 = time.time()
        for p in params.values():
            p.grad = None
        o = O.vid2seq_forward(params, cfg, video, inp, inp != 0, out, out != 0, torch_dropout=tdrop)
        o["loss"].backward()
        with torch.no_grad():
            O.clip_adam_renorm_({k: v.data for k, v in params.items()}, {k: v.grad for k, v in params.items()}, state,
                                lr=3e-4, clip_max_norm=0.1, num_bins=cfg["num_bins"])
        _ = o["loss"].item()
        return time.time() - t0

    t_begin = time.time()
    cands = _pick_threads()
    best = None
    for n in cands:                      # calibration steps double as warm-up
        torch.set_num_threads(n)
        dt = one_step()
        if best is None or dt < best[0]:
            best = (dt, n)
    torch.set_num_threads(best[1])
    times = []
    for i in range(max(0, warmup - len(cands)) + steps):
        dt = one_step()
        if i >= max(0, warmup - len(cands)):
            times.append(dt)
        if time.time() - t_begin > budget_s and len(times) >= 1:
            break
    if not times:
        times = [best[0]]
    times.sort()
    return times[len(times) // 2], len(times), best[1], cands


def cpu_decode_time(cfg, B, T, L, new_tokens, budget_s):
    """Greedy decoding with a KV cache through the oracle port (oracle.greedy_decode_cached) + its encoders, on a
    bounded sample: B videos, `new_tokens` steps."""
    from oracle import vid2seq_oracle as O
    from vidchapters_b200.init import init_state_dict
    sd = init_state_dict(cfg, 0)
    video, inp, _ = synth_batch(B, T, L, 8, 1, cfg["base_vocab"], cfg["base_vocab"] + cfg["num_bins"])
    best = None
    with torch.no_grad():
        for n in _pick_threads():
            torch.set_num_threads(n)
            t0 = time.time()
            ar = O.Arith(False)
            vid = O.vit_forward(sd, cfg, video, ar)
            if cfg["d_model"] != 768:
                vid = ar.linear(vid, sd["proj_v2t.weight"], sd["proj_v2t.bias"])
            enc = O.t5_encoder(sd, cfg, sd["t5_model.shared.weight"][inp], inp != 0, ar)
            memory = torch.cat([vid, enc], 1)
            mask = torch.cat([torch.ones(B, T, dtype=torch.long), (inp != 0).long()], 1)
            ids = O.greedy_decode_cached(sd, cfg, memory, mask, max_new_tokens=new_tokens, stop_when_done=False)
            dt = time.time() - t0
            if best is None or dt < best[0]:
                best = (dt, n, ids.shape[1] - 1)
            if time.time() - t0 > budget_s:
                break
    return best


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation (oracle port of model/vid2seq.py + dvc.py:112-126, fp32) on
    this box's host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = model_cfg(args.model)
    if args.mode == "decode":
        Bs, new = 4, 32
        dt, threads, n_tok = cpu_decode_time(cfg, Bs, T_FRAMES, L_ASR, new, budget_s=120)
        val = Bs * n_tok / dt
        sample = (f"{Bs} videos x {n_tok} new tokens (bounded sample of batch {DECODE_BATCH} x {DECODE_NEW}), encoders + "
                  f"cross K/V + cached greedy loop, torch fp32 CPU, {threads} threads: {dt:.2f} s")
        line = {"impl": "reference", "metric": "greedy_decode_tokens_per_sec", "value": val, "unit": "tokens/s",
                "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"vid2seq {args.model} greedy decode, 100 frames x768, 1000 ASR tok", "sample": sample},
                "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    Bs = min(args.batch or MODEL_BATCH[args.model], REF_VIDEOS)
    med, n, threads, cands = cpu_train_step_time(cfg, Bs, T_FRAMES, L_ASR, S_TGT, args.steps, args.warmup, budget_s=170,
                                                 dropout=args.dropout)
    tokens = Bs * (T_FRAMES + L_ASR + S_TGT)
    val = tokens / med
    sample = (f"{Bs} videos per step (bounded sample of batch {args.batch or MODEL_BATCH[args.model]}: same sequence lengths, "
              f"tokens/s is per-token so the batches compare), dropout {args.dropout} via F.dropout, full dvc.py step via "
              f"the oracle port, torch fp32 CPU; threads calibrated over {cands} -> {threads}; median of {n} steps")
    line = {"impl": "reference", "metric": "train_step_tokens_per_sec", "value": val, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": n, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"vid2seq {args.model} train step (fwd+bwd+clip+adam+renorm), 100 frames x768, 1000 ASR "
                                   "tok, 256 target tok", "sample": sample, "dropout": args.dropout,
                       "videos_per_step": Bs, "host_cores": os.cpu_count()},
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- shared GPU plumbing
class Dist:
    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            torch.distributed.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """EXACTLY `steps` calls bracketed by barrier + synchronize; device time (CUDA events), max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item(), last

    def close(self):
        if self.world > 1:
            torch.distributed.destroy_process_group()


# ----------------------------------------------------------------------------- train arm
def run_train(args):
    args.warmup = max(args.warmup, 3)
    D = Dist()
    dev, world, rank = D.dev, D.world, D.rank
    from vidchapters_b200 import GraphedTrainStep, Vid2Seq, Vid2SeqAdam
    cfg = model_cfg(args.model)
    B, T, L, S = args.batch or MODEL_BATCH[args.model], T_FRAMES, L_ASR, S_TGT
    model = Vid2Seq(args.model, tokenizer=Tok(), vis_drop=args.dropout, enc_drop=args.dropout, dec_drop=args.dropout,
                    seed=0, pretrained=False).to(dev)
    model.train()
    opt = Vid2SeqAdam(model, lr=3e-4, clip_max_norm=0.1, world_size=world)
    ops = model.engine.ops

    video_h, inp_h, out_h = [t.pin_memory() for t in synth_batch(B, T, L, S, 1234 + rank)]
    h2d_bytes = sum(t.numel() * t.element_size() for t in (video_h, inp_h, out_h))
    video_d, inp_d, out_d = video_h.to(dev), inp_h.to(dev), out_h.to(dev)

    def eager_step(host_inputs: bool, read_loss: bool):
        if host_inputs:
            v = video_h.to(dev, non_blocking=True)
            i = inp_h.to(dev, non_blocking=True)
            o = out_h.to(dev, non_blocking=True)
        else:
            v, i, o = video_d, inp_d, out_d
        ld, _ = model(v, {"input_ids": i, "attention_mask": i != 0}, {"input_ids": o, "attention_mask": o != 0})
        opt.zero_grad()
        ld["loss"].backward()
        opt.step()
        return ld["loss"].item() if read_loss else ld["loss"]

    # eager public API (Vid2Seq.forward -> loss.backward() -> Vid2SeqAdam.step()): ~1000 launches/step from Python
    for _ in range(args.warmup):
        eager_step(True, True)
    n_eager = min(args.steps, 3)
    ms_eager, _ = D.timed(lambda: eager_step(True, True), n_eager)
    # graphed public API (GraphedTrainStep): forward+backward replayed as CUDA graph(s), optimiser tail eager
    gstep = GraphedTrainStep(model, opt, video_d, inp_d, out_d, warmup_steps=0)
    sampler = ClockSampler(D.local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        gstep(video_h, inp_h, out_h).item()
    sampler.mark()
    launches0 = ops.launches
    ms_dev, loss_dev = D.timed(lambda: gstep(), args.steps)                                    # inputs resident in HBM
    launches = ops.launches - launches0
    ms_e2e, loss_e2e = D.timed(lambda: gstep(video_h, inp_h, out_h).item(), args.steps)       # host buffers, loss read back
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline: every GEMM launch of one instrumented step, CUDA events on the launch stream
    rec = []
    orig_gemm = ops.gemm

    def timed_gemm(A, Bm, out, **kw):
        a_mn, b_mn = kw.get("a_mn", False), kw.get("b_mn", False)
        M, K = (A.shape[1], A.shape[0]) if a_mn else A.shape
        N = Bm.shape[1] if b_mn else Bm.shape[0]
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig_gemm(A, Bm, out, **kw)
        e.record()
        rec.append((s, e, 2.0 * M * N * K, (M, N, K, int(a_mn), int(b_mn), str(out.dtype).replace("torch.", "") if out is not None else "none",
                                            int(kw.get("act", 0)), int(kw.get("atomic", False)), int(kw.get("splits", 1)))))
        return r

    ops.gemm = timed_gemm
    # the instrumented step runs single-stream: with the visual encoder on its second stream the event pairs of two
    # concurrent kernels overlap and every kernel would be charged the other's time as well
    eng_ = model.engine
    dual_, eng_.dual_stream = eng_.dual_stream, False
    eager_step(False, False)
    torch.cuda.synchronize()
    eng_.dual_stream = dual_
    ops.gemm = orig_gemm
    gemm_ms = sum(r[0].elapsed_time(r[1]) for r in rec)
    gemm_flops = sum(r[2] for r in rec)
    if args.dump_gemms and rank == 0:
        agg = {}
        for s_, e_, f_, key in rec:
            a = agg.setdefault(key, [0, 0.0, 0.0])
            a[0] += 1; a[1] += s_.elapsed_time(e_); a[2] += f_
        os.makedirs(os.path.dirname(args.dump_gemms) or ".", exist_ok=True)
        with open(args.dump_gemms, "w") as fo:
            fo.write("# M N K a_mn b_mn out act atomic splits | launches total_ms TFLOP/s  (one instrumented step)\n")
            for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                fo.write(" ".join(str(x) for x in key) + f" | {a[0]} {a[1]:.3f} {a[2] / (a[1] * 1e-3) / 1e12:.1f}\n")
    pk = measured_peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12

    tokens_step = world * B * (T + L + S)
    ms_step = ms_dev / args.steps
    ms_step_e2e = ms_e2e / args.steps
    if rank != 0:
        D.close()
        return
    step_tflop = FWD_GFLOP_PER_SAMPLE[args.model] * 1e-3 * 3 * B
    traffic = committed_profile("r02_gemm_traffic.json")
    tpipe = committed_profile("r02_tensor_pipe.json") if args.model == "t5-base" and B == 16 else None
    cfg_name = "configs[1]" if args.model == "t5-base" else "configs[3]"
    line = {
        "metric": "train_step_tokens_per_sec", "value": tokens_step / (ms_step * 1e-3), "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"vid2seq {args.model} train step (dvc.py:42-133, generative pass): fwd+bwd+clip+adam+renorm, "
                               f"batch {B}/GPU, 100 frames x768, 1000 ASR tok, 256 target tok (BASELINE {cfg_name})",
                   "global_batch": world * B, "tokens_per_step": tokens_step, "parallelism": f"dp{world}",
                   "dropout": args.dropout, "clip_max_norm": 0.1,
                   "l2": "per-step working set (~6 GB of weights+activations) >> 126 MB L2; no explicit flush",
                   "residual_stream": "fp32", "gemm_operands": "bf16", "accumulate": "fp32",
                   "collectives_per_step": ("gradient all-reduce (flat fp32 buffer, region-wise, overlapped with the backward) "
                                            "with the loss scalar of dvc.py:103 riding in its last region") if world > 1 else "none"},
        "e2e": {"value": tokens_step / (ms_step_e2e * 1e-3), "unit": "tokens/s", "ms_per_step": ms_step_e2e,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "api": "vidchapters_b200.GraphedTrainStep(model, optimizer, ...)(video, input_ids, output_ids).item()"},
        "e2e_eager": {"value": tokens_step / (ms_eager / n_eager * 1e-3), "unit": "tokens/s",
                      "ms_per_step": ms_eager / n_eager,
                      "api": "model(...); optimizer.zero_grad(); loss.backward(); optimizer.step(); loss.item()"},
        "gpu_launches": launches,
        "loss": loss_e2e,
        "step_tflops_algorithmic": step_tflop,
        "step_tensor_frac": (step_tflop / (ms_step * 1e-3)) / pk["tf_sustained"],
        "roofline": {"bound": "tensor", "kernel": "gemm2_bf16_kernel / gemm_bf16_kernel (tcgen05 cta_group::2 / ::1; all %d launches of one step, "
                               "each timed with CUDA events in an eager, single-stream instrumented step)" % len(rec),
                     "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
                     "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step)",
                     "gemm_ms_per_step": gemm_ms, "gemm_share_of_step": gemm_ms / ms_step,
                     "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                     "traffic_source": traffic.get("source") if traffic else None},
        "clocks": clocks,
    }
    if tpipe:
        line["tensor_pipe_pct"] = tpipe
    if not args.no_cpu_baseline:
        Bs = min(B, REF_VIDEOS)
        med, n, threads, cands = cpu_train_step_time(cfg, Bs, T, L, S, steps=3, warmup=2, budget_s=45, dropout=args.dropout)
        line["cpu_baseline"] = {"value": Bs * (T + L + S) / med, "unit": "tokens/s", "cores": threads, "kind": "port",
                                "sample": f"{Bs} videos/step of the same shapes ({T} frames, {L} ASR, {S} target tok), "
                                          f"{args.model} fp32, dropout {args.dropout}; full dvc.py step via the oracle port; "
                                          f"threads calibrated over {cands}; median of {n} steps = {med:.3f} s"}
    print(json.dumps(line), flush=True)
    D.close()


# ----------------------------------------------------------------------------- decode arm (BASELINE configs[4])
def decode_bytes_per_step(cfg, B, E, new_tokens):
    """Algorithmic HBM bytes of ONE greedy step: decoder + LM-head weights (bf16) + cross-attention K/V of
    every layer and sequence + the self-attention K/V written so far (average over the run)."""
    d, dff, nl, V = cfg["d_model"], cfg["d_ff"], cfg["num_layers"], cfg["base_vocab"] + cfg["num_bins"]
    inner = cfg["num_heads"] * cfg["d_kv"]
    w = nl * (4 * d * inner + 4 * d * inner + 2 * d * dff) + V * d
    cross = nl * B * E * 2 * inner
    self_kv = nl * B * (new_tokens / 2) * 2 * inner
    return 2.0 * (w + cross + self_kv)


def run_decode(args):
    D = Dist()
    dev, world, rank = D.dev, D.world, D.rank
    from vidchapters_b200 import Vid2Seq
    cfg = model_cfg(args.model)
    B, T, L, NEW = args.batch or DECODE_BATCH, T_FRAMES, L_ASR, DECODE_NEW
    model = Vid2Seq(args.model, tokenizer=Tok(), seed=0, pretrained=False).to(dev).eval()
    ops = model.engine.ops
    video_h, inp_h, _ = [t.pin_memory() for t in synth_batch(B, T, L, 8, 4321 + rank)]
    h2d_bytes = video_h.numel() * 4 + inp_h.numel() * 8
    video_d, inp_d = video_h.to(dev), inp_h.to(dev)
    tokd = {"input_ids": inp_d, "attention_mask": inp_d != 0}

    def gen_dev():
        model.generate(video_d, tokd, num_beams=1, max_length=NEW)       # random-init weights never emit eos: 256 steps
        return model.last_generated_ids

    def gen_e2e():
        v = video_h.to(dev, non_blocking=True)
        i = inp_h.to(dev, non_blocking=True)
        model.generate(v, {"input_ids": i, "attention_mask": i != 0}, num_beams=1, max_length=NEW)
        return model.last_generated_ids.cpu()

    args.warmup = max(args.warmup, 3)
    steps = min(args.steps, 5)
    sampler = ClockSampler(D.local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        gen_dev()
    sampler.mark()
    launches0 = ops.launches
    ms_dev, ids = D.timed(gen_dev, steps)
    launches = ops.launches - launches0
    ms_e2e, ids_h = D.timed(gen_e2e, steps)
    clocks = sampler.stop() if rank == 0 else None
    n_tok = ids.shape[1] - 1
    # the decode loop alone (what the HBM roofline describes): engine-level timing of the graph replays
    eng = model.engine
    memory, mem_mask, _, E = eng.encode(video_d, inp_d, inp_d != 0)
    torch.cuda.synchronize()
    t_loop = eng.time_greedy_loop(memory, mem_mask, B, E, NEW) if hasattr(eng, "time_greedy_loop") else None
    if rank != 0:
        D.close()
        return
    pk = measured_peaks()
    ms_call = ms_dev / steps
    tokens = world * B * n_tok
    line = {
        "metric": "greedy_decode_tokens_per_sec", "value": tokens / (ms_call * 1e-3), "unit": "tokens/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms_call, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"vid2seq {args.model} greedy decode (Vid2Seq.generate, num_beams=1): batch {B}/GPU, 100 frames "
                               f"x768 + 1000 ASR tok of memory, {n_tok} new tokens (BASELINE configs[4]); one step = one "
                               "generate call (encoders + cross K/V projection + decode loop)",
                   "global_batch": world * B, "new_tokens": n_tok, "parallelism": f"replicas x{world}",
                   "l2": "per-decode-step working set ~3 GB (cross-attention K/V) >> 126 MB L2; no explicit flush"},
        "e2e": {"value": tokens / (ms_e2e / steps * 1e-3), "unit": "tokens/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(ids_h.numel() * 8),
                "api": "vidchapters_b200.Vid2Seq.generate(video, input_tokenized, num_beams=1, max_length=256)"},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if t_loop is not None:
        ms_tok = t_loop / n_tok
        bytes_step = decode_bytes_per_step(cfg, B, E, n_tok)
        ach = bytes_step / (ms_tok * 1e-3) / 1e9
        line["decode_loop"] = {"ms_per_token_step": ms_tok, "tokens_per_sec": B * n_tok / (t_loop * 1e-3)}
        line["roofline"] = {"bound": "hbm", "kernel": "one greedy decode step (CUDA graph: attention over the KV caches + "
                                                      "skinny GEMMs), algorithmic bytes of SURVEY §8d",
                            "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                            "peak_source": pk["source"], "bytes_per_step": bytes_step, "traffic": None}
    if not args.no_cpu_baseline:
        dt, threads, nt = cpu_decode_time(cfg, 2, T, L, 16, budget_s=40)
        line["cpu_baseline"] = {"value": 2 * nt / dt, "unit": "tokens/s", "cores": threads, "kind": "port",
                                "sample": f"2 videos x {nt} new tokens, encoders + cached greedy loop via the oracle port, "
                                          f"torch fp32 CPU: {dt:.2f} s"}
    print(json.dumps(line), flush=True)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="t5-base", choices=["t5-base", "t5-large"])
    ap.add_argument("--mode", default="train", choices=["train", "decode"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: what BASELINE.json quotes for the config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-gemms", default="")
    ap.add_argument("--dropout", type=float, default=0.1, help="vis/enc/dec dropout (reference default 0.1, args.py)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.mode == "decode":
        run_decode(args)
    else:
        run_train(args)


if __name__ == "__main__":
    main()
