"""Multi-GPU check of the data-parallel step (run under torchrun, NCCL): the graphed step with the overlapped,
region-wise all-reduce (GraphedTrainStep, 3 backward phases) must train exactly like the eager step with one flat
all-reduce (Vid2SeqAdam.step), and every rank must end with identical parameters.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check_gpu.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_e2e_gpu import build, GOLD
from vidchapters_b200 import GraphedTrainStep, Vid2SeqAdam

fx = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
cfg = fx["cfg"]
g = torch.Generator().manual_seed(100 + rank)            # a different shard per rank
video = (fx["video"] + 0.1 * torch.randn(fx["video"].shape, generator=g)).cuda()
inp, out = fx["input_ids"].cuda(), fx["output_ids"].cuda()
if rank % 2 == 1:
    inp, out = inp.flip(0), out.flip(0)
it = {"input_ids": inp, "attention_mask": inp != 0}
ot = {"input_ids": out, "attention_mask": out != 0}

m1 = build(cfg); m1.train()
o1 = Vid2SeqAdam(m1, lr=3e-4, clip_max_norm=0.1)
assert o1.world_size == world
for _ in range(4):
    ld, _ = m1(video, it, ot); o1.zero_grad(); ld["loss"].backward(); o1.step()
m2 = build(cfg); m2.train()
o2 = Vid2SeqAdam(m2, lr=3e-4, clip_max_norm=0.1)
ld, _ = m2(video, it, ot); o2.zero_grad(); ld["loss"].backward(); o2.step()
gs = GraphedTrainStep(m2, o2, video, inp, out, warmup_steps=0)
assert gs.graphs, 'data-parallel schedule not active'
for _ in range(3):
    loss = gs(video, inp, out)
torch.cuda.synchronize()
# the returned loss is the rank-mean (dvc.py:103) carried by the gradient all-reduce: compare with an explicit reduction
own = gs.loss.detach().clone().reshape(1)
dist.all_reduce(own)
assert abs(loss.item() - own.item() / world) < 1e-6 * abs(loss.item()), (loss.item(), own.item() / world)
p1, p2 = m1.engine.flat_p, m2.engine.flat_p
err = ((p1 - p2).norm() / p1.norm()).item()
# identical parameters on every rank
ref = p2.clone(); dist.broadcast(ref, 0)
same = torch.equal(ref, p2)
flags = torch.tensor([err, 0.0 if same else 1.0], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"dp_check world={world}: graphed(overlapped all-reduce) vs eager(flat all-reduce) params rel-L2 {flags[0].item():.3e}; "
          f"ranks identical: {flags[1].item() == 0}")
    assert flags[0].item() < 2e-4 and flags[1].item() == 0
    print("DP_CHECK_OK")
dist.destroy_process_group()
