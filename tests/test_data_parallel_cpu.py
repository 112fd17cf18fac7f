"""N>1 host logic on CPU: two gloo ranks, fused optimiser averaging gradients with all-reduce (SURVEY §8e semantics:
'2-rank step == single-process step whose gradient is the average of the per-shard gradients')."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(cfg):
    from oracle.torch_ops import TorchOps
    from vidchapters_b200 import Vid2Seq

    class Tok:
        pad_token_id, eos_token_id = 0, 1

        def __len__(self):
            return cfg["base_vocab"] + cfg["num_bins"]

    return Vid2Seq("t5-base", num_features=cfg["num_features"], depth=cfg["depth"], tokenizer=Tok(), dec_drop=0.0,
                   t5_config=cfg, ops=TorchOps())


def _batch(rank):
    g = torch.Generator().manual_seed(100 + rank)
    video = torch.randn(1, 6, 768, generator=g)
    inp = torch.randint(2, 1100, (1, 12), generator=g)
    out = torch.randint(2, 1100, (1, 8), generator=g)
    if rank == 1:
        out[0, -3:] = 0
    return video, {"input_ids": inp, "attention_mask": inp != 0}, {"input_ids": out, "attention_mask": out != 0}


def _cfg():
    from vidchapters_b200 import TINY
    return dict(TINY, num_features=6, depth=1, num_layers=1)


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from vidchapters_b200 import Vid2SeqAdam
    m = _make(_cfg())
    opt = Vid2SeqAdam(m, lr=3e-4, clip_max_norm=0.1)
    assert opt.world_size == 2
    v, it, ot = _batch(rank)
    ld, _ = m(v, it, ot)
    opt.zero_grad()
    ld["loss"].backward()
    opt.step()
    torch.save({n: p.detach().clone() for n, p in m._params.items()}, os.path.join(outdir, f"rank{rank}.pt"))
    # the graphed step's schedule without the graphs: the backward phases of engine.dp_phases(), each finished region of the flat gradient
    # buffer all-reduced asynchronously while the next phase runs (vidchapters_b200/graphed.py)
    m2 = _make(_cfg())
    opt2 = Vid2SeqAdam(m2, lr=3e-4, clip_max_norm=0.1)
    eng = m2.engine
    m2._refresh_shadow()          # bf16 shadow of the freshly initialised fp32 parameters (the module's forward does this)
    loss, ectx = eng.forward(v, it["input_ids"], it["attention_mask"], ot["input_ids"], ot["attention_mask"], training=True)
    eng.zero_grad()
    works = []
    for ph, regions in eng.dp_phases():
        eng.backward(ectx, phase=ph)
        for lo, hi in regions:
            works.append(torch.distributed.all_reduce(eng.flat_g[lo:hi], async_op=True))
    for w in works:
        w.wait()
    m2._end_backward()
    opt2.step(grads_already_reduced=True)
    torch.save({n: p.detach().clone() for n, p in m2._params.items()}, os.path.join(outdir, f"rank{rank}_phased.pt"))
    torch.distributed.destroy_process_group()


def test_two_rank_step_equals_averaged_gradient_step(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    p0 = torch.load(tmp_path / "rank0_phased.pt")
    p1 = torch.load(tmp_path / "rank1_phased.pt")
    for n in r0:
        assert torch.equal(r0[n], r1[n]), n      # replicas stay identical after the all-reduced step
        assert torch.equal(p0[n], p1[n]), n
        assert torch.equal(p0[n], r0[n]), n      # region-wise overlapped all-reduce == one flat all-reduce, bit for bit
    # single process: average of the two shard gradients, then the same tail
    from vidchapters_b200 import Vid2SeqAdam
    torch.set_num_threads(2)  # same thread count as the workers: identical CPU matmul blocking, hence identical
    # bf16 rounding decisions in the emulated kernels (a 1-ulp fp32 difference can flip a bf16 rounding)
    m = _make(_cfg())
    before = {n: p.detach().clone() for n, p in m._params.items()}
    eng = m.engine
    acc = None
    for rank in (0, 1):
        v, it, ot = _batch(rank)
        ld, _ = m(v, it, ot)
        for p in m.parameters():
            p.grad = None
        ld["loss"].backward()
        acc = eng.flat_g.clone() if acc is None else acc + eng.flat_g
    eng.flat_g.copy_(acc / 2)
    Vid2SeqAdam(m, lr=3e-4, clip_max_norm=0.1, world_size=1).step()
    for n, p in m._params.items():
        d_ref, d = p.detach() - before[n], r0[n] - before[n]
        assert ((d - d_ref).norm() / (d_ref.norm() + 1e-30)).item() < 2e-3, n
