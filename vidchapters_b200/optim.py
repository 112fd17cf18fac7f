"""Fused optimiser for vidchapters_b200.Vid2Seq with the torch.optim surface dvc.py uses.

Reference sequence being replaced (dvc.py:112-126,345-351; util/misc.py:15-42 writes `param_groups[0]["lr"]`):
    optimizer.zero_grad(); loss.backward(); clip_grad_norm_(params, max_norm); optimizer.step(); <time-token renorm>
Here `step()` runs, over the model's flat buffers: [data-parallel gradient all-reduce, NCCL] -> global grad norm ->
clip + Adam + bf16 weight re-pack (one kernel) -> time-token renorm.  The reference never all-reduces gradients
(its ranks silently diverge); with world_size > 1 this optimiser averages them, i.e. standard DDP.
"""
from __future__ import annotations

import torch


class Vid2SeqAdam:
    def __init__(self, model, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_max_norm=1.0,
                 renorm_time_tokens=True, process_group=None, world_size=None):
        if weight_decay != 0.0:
            raise NotImplementedError("the reference trains with weight_decay 0 (dvc.py:345-351)")
        self.model = model
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=0.0, params=list(model.parameters()))]
        self.clip_max_norm = clip_max_norm
        self.renorm = renorm_time_tokens
        self.pg = process_group
        if world_size is None:
            world_size = torch.distributed.get_world_size(process_group) if (
                torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        self.world_size = world_size
        self.last_grad_norm_sq = None

    def zero_grad(self, set_to_none: bool = True):
        for p in self.param_groups[0]["params"]:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self, grads_already_reduced: bool = False):
        eng = self.model.engine
        g = self.param_groups[0]
        grad_scale = 1.0
        if self.world_size > 1:
            if not grads_already_reduced:   # (GraphedTrainStep overlaps the all-reduce with the backward itself)
                torch.distributed.all_reduce(eng.flat_g, group=self.pg)  # SUM over ranks; averaged by grad_scale below
            grad_scale = 1.0 / self.world_size
        self.last_grad_norm_sq = eng.optimizer_step(g["lr"], betas=g["betas"], eps=g["eps"],
                                                    clip_max_norm=self.clip_max_norm, grad_scale=grad_scale,
                                                    renorm=self.renorm)
        self.model._shadow_valid = True

    # checkpoint / resume (dvc.py:402-441 saves optimizer.state_dict())
    def state_dict(self):
        eng = self.model.engine
        return {"step": eng.adam_step_count, "exp_avg": eng.adam_m, "exp_avg_sq": eng.adam_v,
                "lr": self.param_groups[0]["lr"]}

    def load_state_dict(self, sd):
        eng = self.model.engine
        eng.adam_step_count = int(sd["step"])
        if sd["exp_avg"] is not None:
            eng.adam_m = sd["exp_avg"].to(eng.device).clone()
            eng.adam_v = sd["exp_avg_sq"].to(eng.device).clone()
        self.param_groups[0]["lr"] = sd.get("lr", self.param_groups[0]["lr"])
