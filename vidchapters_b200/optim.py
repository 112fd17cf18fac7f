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
        from .vid2seq import _dev_guard
        with _dev_guard(self.model._flat.device):
            self._step(grads_already_reduced)

    def _step(self, grads_already_reduced: bool = False):
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

    # checkpoint / resume: dvc.py:402-441 saves `optimizer.state_dict()` and dvc.py:359-361 (--resume) feeds it back to
    # `optimizer.load_state_dict`.  The format is torch.optim.Adam's own — {"state": {i: {"step", "exp_avg",
    # "exp_avg_sq"}}, "param_groups": [{..., "params": [0..n-1]}]} with i indexing `model.parameters()` (tied weights
    # once), which is the order of the reference model's parameters as well (tests/test_host_logic_cpu.py) — so a
    # checkpoint written by the reference resumes here and vice versa.  The flat moment buffers are only views' backing.
    def _param_names(self):
        return [n for n, _ in self.model.named_parameters()]

    def state_dict(self):
        eng = self.model.engine
        names = self._param_names()
        state = {}
        if eng.adam_m is not None:
            for i, n in enumerate(names):
                o, shp, k = eng.layout[n]
                state[i] = {"step": torch.tensor(float(eng.adam_step_count)),
                            "exp_avg": eng.adam_m[o:o + k].view(shp), "exp_avg_sq": eng.adam_v[o:o + k].view(shp)}
        g = self.param_groups[0]
        group = dict(lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"], weight_decay=0.0, amsgrad=False, maximize=False,
                     foreach=None, capturable=False, differentiable=False, fused=None,
                     params=list(range(len(names))), param_names=names)
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        eng = self.model.engine
        names = self._param_names()
        if "state" not in sd or "param_groups" not in sd:
            raise ValueError("Vid2SeqAdam.load_state_dict expects torch.optim.Adam's format {'state', 'param_groups'} "
                             "(what the reference's dvc.py checkpoints hold)")
        group = sd["param_groups"][0]
        if len(group["params"]) != len(names):
            raise ValueError(f"optimizer state has {len(group['params'])} parameters, the model has {len(names)}")
        if "param_names" in group and list(group["param_names"]) != names:
            raise ValueError("optimizer state was saved for a different parameter order / model")
        if group.get("amsgrad") or group.get("weight_decay", 0.0) != 0.0 or group.get("maximize"):
            raise NotImplementedError("only plain Adam (weight_decay 0, no amsgrad), as dvc.py:345-351 builds it")
        g = self.param_groups[0]
        g["lr"], g["betas"], g["eps"] = group["lr"], tuple(group["betas"]), group["eps"]
        state = sd["state"]
        if not state:
            eng.adam_m = eng.adam_v = None
            eng.adam_step_count = 0
            return
        if eng.adam_m is None:
            eng.adam_m = torch.zeros_like(eng.flat_p)
            eng.adam_v = torch.zeros_like(eng.flat_p)
        steps = set()
        for pid, n in zip(group["params"], names):
            st = state[pid] if pid in state else state[str(pid)]
            o, shp, k = eng.layout[n]
            if tuple(st["exp_avg"].shape) != tuple(shp):
                raise ValueError(f"{n}: moment shape {tuple(st['exp_avg'].shape)} != parameter shape {tuple(shp)}")
            eng.adam_m[o:o + k].copy_(st["exp_avg"].reshape(-1))
            eng.adam_v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(st["step"]))
        if len(steps) != 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): not a dvc.py Adam state")
        eng.adam_step_count = steps.pop()
