"""Static-shape train step replayed as a CUDA graph.

A dvc.py step at fixed (B, T, L, S) is ~1000 kernel launches; issued one by one from Python the host cannot keep a
B200 busy (measured: 75 ms/step of which the GPU works ~45 ms).  `GraphedTrainStep` captures forward + backward
(zeroing of the flat gradient buffer included) once into a `torch.cuda.CUDAGraph` and replays it per step; the
optimiser tail (gradient all-reduce for N>1, clip + Adam + renorm: 5 launches) stays eager so NCCL is never captured
and the learning rate / Adam step count remain plain host values.  The engine's second stream (visual encoder next to
the text encoder) is forked and joined inside the capture, i.e. becomes parallel branches of the graph.  For N>1 the
backward is captured as three graphs and the all-reduce of each finished gradient region overlaps the next one.

The semantics are exactly `loss_dict, _ = model(...); optimizer.zero_grad(); loss.backward(); optimizer.step()`
(dvc.py:70-116) on the batch copied into the static input buffers.
"""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, video, input_ids, output_ids, warmup_steps: int = 2):
        from .vid2seq import _dev_guard
        with _dev_guard(model._flat.device):
            self._init(model, optimizer, video, input_ids, output_ids, warmup_steps)

    def _init(self, model, optimizer, video, input_ids, output_ids, warmup_steps):
        """`video` (B,T,768) float, `input_ids` (B,L), `output_ids` (B,S) int64: example batch (any device; shapes are
        frozen).  Masks are `ids != 0` as in dvc.py:44-53.  NOTE: the warm-up runs real optimiser steps on the example
        batch unless warmup_steps=0 (then the caller must have run at least one eager step at these shapes)."""
        self.model, self.optimizer = model, optimizer
        dev = model._flat.device
        assert dev.type == "cuda", "GraphedTrainStep needs the model on a CUDA device"
        self.video = video.to(dev, copy=True).float().contiguous()
        self.input_ids = input_ids.to(dev, copy=True).contiguous()
        self.output_ids = output_ids.to(dev, copy=True).contiguous()
        self.graph = torch.cuda.CUDAGraph()
        # dropout: the captured kernels carry fixed seeds; a device-side salt, bumped before every replay, gives each
        # step its own masks (forward and backward of one step read the same salt value)
        self.salt = torch.zeros(1, dtype=torch.int32, device=dev)
        model.engine.ops.set_dropout_salt(self.salt)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup_steps):
                self._fwd_bwd()
                optimizer.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        model._refresh_shadow()            # outside the graph: the fused optimiser keeps the shadow valid afterwards
        model._shadow_valid = True
        ops = model.engine.ops
        n0 = ops.launches
        self.world = getattr(optimizer, "world_size", 1)
        self.graphs = []          # data parallel: one graph per backward phase after the first
        self.phases = None
        if self.world > 1:
            # data parallel: the backward is captured as the phases of engine.dp_phases() — forward + LM head + decoder |
            # visual encoder next to the last third of the text encoder | middle third | first third + embeddings — and
            # NCCL all-reduces each finished region of the gradient buffer while the following phases run
            eng = model.engine
            self.phases = eng.dp_phases()
            with torch.cuda.graph(self.graph):
                self.loss = self._fwd_bwd(phase=0)
            for k, (ph, _) in enumerate(self.phases[1:]):
                g = torch.cuda.CUDAGraph()
                last = k == len(self.phases) - 2
                with torch.cuda.graph(g, pool=self.graph.pool()):
                    eng.backward(self._ectx, phase=ph)
                    if last:
                        self._ectx = None
                        self.model._end_backward()
                self.graphs.append(g)
            self.loss_avg = torch.zeros((), dtype=torch.float32, device=dev)
        else:
            with torch.cuda.graph(self.graph):
                self.loss = self._fwd_bwd()
        self.launches_per_replay = ops.launches - n0   # libvidchap kernels inside the captured graph(s)
        self.replays = 0

    def close(self):
        """Detaches the dropout salt from the library (the captured graphs keep reading `self.salt`, so the object must
        stay alive while it is replayed; after close() eager launches draw un-salted masks again)."""
        try:
            if getattr(self.model.engine.ops, "_salt", None) is self.salt:
                self.model.engine.ops.set_dropout_salt(None)
        except Exception:
            pass

    def __del__(self):
        self.close()

    def _fwd_bwd(self, phase=None):
        """Forward + backward through the engine directly (no autograd engine inside the capture: its cross-stream
        bookkeeping for leaf tensors is not capture-safe); gradients land in the flat buffer that `p.grad` views."""
        m = self.model
        eng = m.engine
        eng.drop_rates = dict(vis=m.vis_drop, enc=m.enc_drop, dec=m.dec_drop)
        ids, out = self.input_ids, self.output_ids
        eng.zero_grad()            # before the forward: the 1.2 GB fill hides under the two-stream encoder phase
        loss, ectx = eng.forward(self.video, ids, ids != 0, out, out != 0, training=m.training)
        eng.backward(ectx, phase=phase)
        if phase == 0:
            eng.loss_slot.copy_(loss.view(1))   # rides through the all-reduce of the last gradient region (dvc.py:103)
            self._ectx = ectx                   # the later phases are captured into their own graphs
        else:
            m._end_backward()      # host-side: make every Parameter's .grad a view of the flat gradient buffer
        return loss.view(())

    def __call__(self, video=None, input_ids=None, output_ids=None):
        """Copies the batch (host pinned or device tensors) into the static buffers, replays forward+backward, runs the
        optimiser tail.  Returns the (static) 0-dim device loss tensor; `.item()` it to read the value."""
        from .vid2seq import _dev_guard
        with _dev_guard(self.model._flat.device):
            return self._call(video, input_ids, output_ids)

    def _call(self, video=None, input_ids=None, output_ids=None):
        if video is not None:
            self.video.copy_(video, non_blocking=True)
        if input_ids is not None:
            self.input_ids.copy_(input_ids, non_blocking=True)
        if output_ids is not None:
            self.output_ids.copy_(output_ids, non_blocking=True)
        if not self.model._shadow_valid:     # parameters were changed from outside (load_state_dict, stock optimiser)
            self.model.engine.sync_bf16()
        self.salt.add_(0x9E3779B)   # new dropout masks for this step
        self.graph.replay()
        self.replays += 1
        self.model.engine.ops.launches += self.launches_per_replay
        if not self.graphs:
            self.optimizer.step()
            return self.loss
        eng = self.model.engine
        pg = self.optimizer.pg
        # each region of flat_g is all-reduced by NCCL (its own stream) as soon as the phase that finishes it has been
        # enqueued, overlapping the phases that follow; the last region starts at the loss slot, so the rank-mean loss of
        # dvc.py:103 (util/dist.py:89-113) costs no collective of its own
        works = []
        n_ph = len(self.phases)
        for k, (ph, regions) in enumerate(self.phases):
            if k > 0:
                self.graphs[k - 1].replay()
            for r, (lo, hi) in enumerate(regions):
                if k == n_ph - 1 and r == len(regions) - 1:
                    assert lo == 0
                    buf = eng._g_store[60:64 + hi]          # [.., loss slot | shared | first encoder layers]
                else:
                    buf = eng.flat_g[lo:hi]
                works.append(torch.distributed.all_reduce(buf, group=pg, async_op=True))
        for w in works:
            w.wait()
        self.optimizer.step(grads_already_reduced=True)
        torch.div(eng.loss_slot[0], float(self.world), out=self.loss_avg)
        return self.loss_avg
