"""The fused semi-supervised training step -- what the reference's ``train()`` loop body does
(main_ucf101.py:169-184: zero_grad, train_model_interface :50-150, loss.backward(), Adam step) -- scheduled
for B200:

  * both forward passes (clips and their horizontal flips) run as ONE batch of 2P clips; BatchNorm keeps the
    reference's per-pass statistics through the kernels' `groups` argument (STATE.bn_groups = 2);
  * losses, consistency masks and their gradients stay on the device (no numpy / host round trip);
  * parameters, gradients and Adam moments live in flat fp32 buffers: one fused Adam launch, one (bucketed)
    NCCL all-reduce, gradients accumulated by the kernels straight into the flat buffer;
  * no host synchronisation inside the step (losses are returned as device tensors).

Host logic only; every number is produced by libb200caps.so kernels (torch is the allocator / autograd tape).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import engine, ops
from .plans import View, act_dtype


@dataclass
class StepArgs:
    """The flags of main_ucf101.py:283-318 that enter the step."""
    bv: bool = True
    gv: bool = False
    n_frames: int = 5
    predict_maps: bool = False
    bv_wt: float = 0.5
    gv_wt: float = 0.5
    lower_thresh: Optional[float] = None
    upper_thresh: Optional[float] = None
    wt_loc: float = 1.0
    wt_cls: float = 1.0
    wt_cons: float = 0.1
    thresh_epoch: int = 11
    lr: float = 1e-4
    rampup_epochs: int = 100      # exp_rampup(N_EPOCHS), main_ucf101.py:419


def exp_rampup(rampup_length: int, epoch: float) -> float:
    if epoch < rampup_length:
        e = float(np.clip(epoch, 0.0, rampup_length))
        phase = 1.0 - e / rampup_length
        return float(np.exp(-5.0 * phase * phase))
    return 1.0


class FlatParams:
    """Re-homes every parameter of `model` (and its gradient) as a view into one flat fp32 buffer, in
    registration order.  state_dict keys / shapes are unchanged."""

    def __init__(self, model: torch.nn.Module):
        params = [p for p in model.parameters()]
        dev = params[0].device
        offs, n = [], 0
        for p in params:
            n = (n + 3) // 4 * 4          # 16-byte alignment of every tensor
            offs.append(n)
            n += p.numel()
        n = (n + 3) // 4 * 4
        self.n = n
        self.data = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.offsets: Dict[str, int] = {}
        names = {id(p): k for k, p in model.named_parameters()}
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.data[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
                self.offsets[names[id(p)]] = o
        engine.bump_weights_epoch()

    def zero_grad(self):
        ops.fill_f32(self.grad, 0.0)


class TrainStep:
    """Eager:   step = TrainStep(model, args); out = step(data, fl_data, action, seg, labels_host)
    Graphed: step.capture(P, labels_host); out = step.replay(data, fl_data, action, seg)  (one cudaGraphLaunch per step)"""

    def __init__(self, model: torch.nn.Module, args: StepArgs = StepArgs(), process_group=None):
        self.model = model
        self.args = args
        self.flat = FlatParams(model)
        dev = self.flat.data.device
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)    # Adam step counter (device: graph replayable)
        from .ddp import GradBuckets, bucket_ranges
        named = dict(model.named_parameters())
        ranges = bucket_ranges(self.flat.offsets, {k: p.numel() for k, p in named.items()}, self.flat.n)
        self.buckets = GradBuckets(self.flat.grad, ranges, process_group)
        self.world = self.buckets.world
        self.rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
        self._steps_seen = 0
        self.arena = None
        self.bn_counters = None
        self._staging = None
        self._staged = False
        self.packs = ops.PackRegistry()
        self.graph = None
        self.graph_opt = None
        self.static = None
        self.launches_per_step = None
        # Schedule scalars the kernels read from DEVICE memory, so one captured graph serves every epoch / learning rate:
        # [0] lr  [1] wt_ramp [2] bv_wt [3] gv_wt  [4] a_l2 [5] a_lv [6] a_lg (consistency-gradient weights x wt_cons)
        # [7] 1.0 once epoch >= thresh_epoch (unlabeled clips are pose-masked with their arg-max class, capsules_ucf101.py:466)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self._schedule = None
        self.set_schedule(epoch=1, lr=args.lr)

    def set_schedule(self, epoch: Optional[float] = None, lr: Optional[float] = None):
        """Per-epoch host scalars of the reference's loop -- wt_ramp = exp_rampup(N_EPOCHS)(epoch) (main_ucf101.py:181,419),
        the learning rate after ReduceLROnPlateau (:417,456), the epoch >= thresh_epoch switch -- uploaded to the device
        vector the kernels read.  Cheap (one 32-byte copy); call it whenever either changes, also between graph replays."""
        a = self.args
        ep = self._schedule[0] if (epoch is None and self._schedule) else (1 if epoch is None else epoch)
        lr = self._schedule[1] if (lr is None and self._schedule) else (a.lr if lr is None else lr)
        if self._schedule == (ep, lr):
            return
        wt_ramp = exp_rampup(a.rampup_epochs, ep)
        mode = (1 if a.bv else 0) | (2 if a.gv else 0)
        if mode == 3:
            a_l2, a_lv, a_lg = a.bv_wt * (1 - wt_ramp), a.bv_wt * wt_ramp, a.gv_wt
        elif mode == 2:
            a_l2, a_lv, a_lg = 0.0, 0.0, 1.0
        elif mode == 1:
            a_l2, a_lv, a_lg = 1 - wt_ramp, wt_ramp, 0.0
        else:
            a_l2, a_lv, a_lg = 1.0, 0.0, 0.0
        vals = [lr, wt_ramp, a.bv_wt, a.gv_wt, a_l2 * a.wt_cons, a_lv * a.wt_cons, a_lg * a.wt_cons,
                1.0 if ep >= a.thresh_epoch else 0.0]
        self.hyper.copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=False)
        self._schedule = (ep, lr)

    @staticmethod
    def _label_tensors(labels_host, dev):
        labels_host = torch.as_tensor(labels_host).float().cpu()
        lab_idx = torch.nonzero(labels_host == 1).view(-1).to(torch.int32)
        return lab_idx.to(dev), labels_host.to(dev), int(lab_idx.numel())

    def __call__(self, data, fl_data, action, seg, labels_host, epoch: int = 1):
        """data / fl_data (P,3,8,H,W) fp32 CUDA, action (P,1) CUDA, seg (P,1,8,H,W) fp32 CUDA, labels_host: CPU tensor /
        list with 1 = labeled.  Returns dict of device scalars (total, loc, cls, cons) and the step's outputs.
        uint8 input pipeline: data (P,3,8,H,W) uint8 as decoded + fl_data=None (+ optionally a uint8 seg): the /255
        scaling and the mirrored second pass are produced on the device (ucf_dataloader.py:162-185)."""
        engine.require_cuda(data, "data")
        lab_idx, labels_dev, n_lab = self._label_tensors(labels_host, data.device)
        self.set_schedule(epoch=epoch)
        return self._impl(data, fl_data, action, seg, lab_idx, labels_dev, n_lab)

    def _optimizer(self):
        a, flat = self.args, self.flat
        ops.adam_step(flat.data, flat.grad, flat.m, flat.v, flat.n, a.lr, 0.9, 0.999, 1e-6, self.step_dev, 1.0 / self.world,
                      lr_dev=self.hyper[0:1])
        engine.bump_weights_epoch()

    def _snapshot(self):
        """Everything a training step mutates: weights, Adam moments + step counter, BatchNorm buffers."""
        bufs = list(self.model.buffers())
        return dict(data=self.flat.data.clone(), m=self.flat.m.clone(), v=self.flat.v.clone(), step=self.step_dev.clone(),
                    bufs=[b.clone() for b in bufs])

    def _restore(self, snap):
        with torch.no_grad():
            self.flat.data.copy_(snap["data"])
            self.flat.m.copy_(snap["m"])
            self.flat.v.copy_(snap["v"])
            self.step_dev.copy_(snap["step"])
            for b, saved in zip(self.model.buffers(), snap["bufs"]):
                b.copy_(saved)
        engine.bump_weights_epoch()

    # ---- CUDA graph ------------------------------------------------------------------------------------
    def capture(self, P: int, labels_host, epoch: int = 1, T: int = 8, H: int = 224, W: int = 224, warmup: int = 3,
                init_batch=None, ddp_graph: Optional[str] = None, uint8_inputs: bool = False):
        """Capture the whole step (both passes, losses, backward, all-reduce, Adam, weight re-packing) into one CUDA
        graph with static input buffers.  Only the labeled/unlabeled PATTERN is baked into the graph; the epoch-dependent
        scalars and the learning rate are read from device memory (set_schedule), so one capture serves the whole run.
        `warmup` uncaptured steps run first (they size the allocator pools, build the kernel plans and register the
        weight-packing jobs); the model weights, Adam state and BatchNorm buffers are snapshotted before and restored
        after them, so capture() leaves the training state exactly as it found it."""
        dev = self.flat.data.device
        if uint8_inputs:
            # static inputs = what the dataloader decodes: uint8 clips and masks; fl_data does not exist on the host side
            st = dict(data=torch.randint(0, 256, (P, 3, T, H, W), dtype=torch.uint8, device=dev), fl_data=None,
                      action=torch.zeros((P, 1), device=dev), seg=torch.zeros((P, 1, T, H, W), dtype=torch.uint8, device=dev))
        else:
            st = dict(data=torch.rand((P, 3, T, H, W), device=dev), fl_data=torch.rand((P, 3, T, H, W), device=dev),
                      action=torch.zeros((P, 1), device=dev), seg=torch.zeros((P, 1, T, H, W), device=dev))
        if init_batch is not None:
            for k, v in zip(("data", "fl_data", "action", "seg"), init_batch):
                if st[k] is not None:
                    st[k].copy_(v)
        lab_idx, labels_dev, n_lab = self._label_tensors(labels_host, dev)
        self.set_schedule(epoch=epoch)
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):      # plans, packed-weight buffers, kernel attributes, allocator pools, NCCL channels
                self._impl(st["data"], st["fl_data"], st["action"], st["seg"], lab_idx, labels_dev, n_lab)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap)
        del snap
        from . import _abi
        l0 = _abi.launch_count()
        g = torch.cuda.CUDAGraph()
        mode = ddp_graph or os.environ.get("B2C_DDP_GRAPH", "split")   # "single" works but hangs NCCL teardown at exit
        self.graph_opt = None
        if self.world == 1 and mode != "split":
            with torch.cuda.graph(g):
                out = self._impl(st["data"], st["fl_data"], st["action"], st["seg"], lab_idx, labels_dev, n_lab)
        elif mode == "single":
            # data parallel, one graph: the two bucketed NCCL all-reduces are captured on the communication stream
            # (forked from / joined to the capture stream by events), the first one under the encoder backward
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                out = self._impl(st["data"], st["fl_data"], st["action"], st["seg"], lab_idx, labels_dev, n_lab)
        else:
            # "split": graph 1 = forward + backward, ONE eager NCCL all-reduce of the flat gradient buffer,
            # graph 2 = Adam (weight re-packing happens at the start of graph 1 of the next step).  The split is a
            # property of THIS capture only: eager calls made later still all-reduce and step the optimiser.
            with torch.cuda.graph(g):
                out = self._impl(st["data"], st["fl_data"], st["action"], st["seg"], lab_idx, labels_dev, n_lab, split=True)
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt):
                self._optimizer()
        self.launches_per_step = _abi.launch_count() - l0
        self.graph, self.static, self.static_out = g, st, out
        # The graph reads these by address and they live OUTSIDE its private pool, so they must outlive this call
        # (dropping them lets later allocations reuse the memory -> garbage labeled-clip indices inside the graph).
        self._graph_refs = (lab_idx, labels_dev, side)
        return self

    def prefetch(self, data, fl_data, action, seg):
        """Start copying the NEXT step's (pinned host) inputs to the device on a copy stream, into a staging set, while
        the current step computes.  The next replay() called without inputs moves them into the graph's static buffers
        (device to device, ~0.1 ms) before it launches: the 180 MB/step H2D transfer leaves the critical path."""
        if self._staging is None:
            self._staging = {k: torch.empty_like(v) for k, v in self.static.items() if v is not None}
            self._copy_stream = torch.cuda.Stream()
            self._staging_ready = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
        cs = self._copy_stream
        cs.wait_event(self._staging_free)          # the previous step has drained the staging set
        with torch.cuda.stream(cs):
            for k, v in (("data", data), ("fl_data", fl_data), ("action", action), ("seg", seg)):
                if k in self._staging:
                    self._staging[k].copy_(v, non_blocking=True)
            self._staging_ready.record(cs)
        self._staged = True

    def replay(self, data=None, fl_data=None, action=None, seg=None, epoch=None, lr=None):
        """Copy the (host or device) inputs into the static buffers on the current stream and launch the graph.
        Without arguments: use the inputs staged by prefetch(), else whatever the static buffers hold.
        epoch / lr: new schedule point (see set_schedule) -- no re-capture needed."""
        st = self.static
        if epoch is not None or lr is not None:
            self.set_schedule(epoch=epoch, lr=lr)
        if data is None and self._staged:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._staging_ready)
            for k in self._staging:
                st[k].copy_(self._staging[k], non_blocking=True)
            self._staging_free.record(cur)
            self._staged = False
        for k, v in (("data", data), ("fl_data", fl_data), ("action", action), ("seg", seg)):
            if v is not None and st[k] is not None:
                st[k].copy_(v, non_blocking=True)
        self.graph.replay()
        if self.graph_opt is not None:
            if self.world > 1:
                dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM, group=self.buckets.group)
            self.graph_opt.replay()
        # the graph updated the fp32 weights through raw pointers (tensor._version did not move): packed operands that
        # the module path (eval / validation forward) caches are stale now
        engine.bump_weights_epoch()
        return self.static_out

    # ---- the step ------------------------------------------------------------------------------------------
    def _impl(self, data, fl_data, action, seg, lab_idx, labels_dev, n_lab: int, split: bool = False):
        """Hand-scheduled forward + backward: the model topology is fixed, so the step walks it explicitly and calls
        the forward / backward halves of the engine functions directly (no autograd engine, no tape threads) -- which
        also makes the whole step capturable in one CUDA graph.  split=True (graph capture on > 1 GPU only): stop after
        the backward pass; the caller all-reduces and runs the optimiser."""
        with torch.no_grad():
            return self._impl_nograd(data, fl_data, action, seg, lab_idx, labels_dev, n_lab, split)

    def _impl_nograd(self, data, fl_data, action, seg, lab_idx, labels_dev, n_lab: int, split: bool):
        a, model, flat = self.args, self.model, self.flat
        hyper = self.hyper
        E = engine
        P = data.shape[0]
        H, W = data.shape[-2], data.shape[-1]
        dev = data.device
        C = model.NUM_CLASSES
        model.train()
        flat.zero_grad()
        if self.arena is None:
            self.arena = E.ZeroArena(1 << 18, dev)       # 1 MB of fp32: BatchNorm sum workspaces of one step (~90 x <= 3 KB)
            self.bn_counters = [m.num_batches_tracked for m in model.modules() if isinstance(m, torch.nn.BatchNorm3d)]
        self.arena.reset()
        E.STATE.zero_arena = self.arena
        E.STATE.defer_bn_counters = True
        # all derived bf16 operand tiles in one launch (after the first step every packing job is registered)
        ops.PACKS = self.packs
        if self.packs.flushed_epoch != E.STATE.weights_epoch and self._steps_seen >= 1:
            self.packs.flush(E.STATE.weights_epoch)
        self._steps_seen += 1
        E.STATE.direct_grads = True
        E.STATE.bn_groups = 2
        tape = []   # (backward closure) in forward order
        try:
            # ---------------- forward: both passes as one batch [clips ; flipped clips] ----------------
            i3d = model.conv1
            Tn = data.shape[2]
            stem = i3d.Conv3d_1a_7x7._layer
            prefolded = None
            if isinstance(stem, E.StemLayer) and stem.use_fold((Tn, H, W)) and data.shape[1] <= 4:
                # clips go straight into the stem's folded layout (no channels-last copy of the input)
                pt = stem.geometry((Tn, H, W))[1][0]
                x_cl = torch.empty((2 * P, 1, H, W, stem.KF), dtype=act_dtype(), device=dev)
                if data.dtype == torch.uint8:
                    assert fl_data is None, "uint8 input pipeline: the mirrored pass is produced on the device"
                    ops.clips_to_folded(data.contiguous(), x_cl, pt, stem.KF // 4, mirror=True)
                else:
                    ops.clips_to_folded(data.contiguous(), x_cl, pt, stem.KF // 4, mirror=False)
                    ops.clips_to_folded(fl_data.contiguous(), x_cl[P:], pt, stem.KF // 4, mirror=False)
                prefolded = (Tn, H, W)
            else:
                x_cl = torch.empty((2 * P, Tn, H, W, 8), dtype=act_dtype(), device=dev)
                if data.dtype == torch.uint8:
                    assert fl_data is None, "uint8 input pipeline: the mirrored pass is produced on the device"
                    ops.u8_clip_to_cl(data.contiguous(), x_cl)
                else:
                    ops.ncdhw_to_cl(data.contiguous(), 8, out=x_cl[:P])
                    ops.ncdhw_to_cl(fl_data.contiguous(), 8, out=x_cl[P:])

            def unit(mod, x, need_dx=True, prefolded=None):
                ctx = _Ctx((need_dx, True, True, True, False))
                ctx.prefolded = prefolded
                y = E.Unit3DFn.forward(ctx, x, mod.conv3d.weight, mod.bn.weight, mod.bn.bias, mod)
                return y, (lambda g, ctx=ctx: E.Unit3DFn.backward(ctx, g)[0])

            def pool(mod, x):
                ctx = _Ctx((True, False, False))
                y = E.MaxPoolFn.forward(ctx, x, tuple(mod.kernel_size), tuple(mod.stride))
                return y, (lambda g, ctx=ctx: E.MaxPoolFn.backward(ctx, g)[0])

            def incep(mod, x):
                ctx = _Ctx((True,) + (False,) * 19)
                params = []
                for n in E.InceptionFn.UNITS:
                    u = getattr(mod, n)
                    params += [u.conv3d.weight, u.bn.weight, u.bn.bias]
                y = E.InceptionFn.forward(ctx, x, mod, *params)
                return y, (lambda g, ctx=ctx: E.InceptionFn.backward(ctx, g)[0])

            y1, b_stem = unit(i3d.Conv3d_1a_7x7, x_cl, need_dx=False, prefolded=prefolded)          # cross112
            p1, b_p1 = pool(i3d.MaxPool3d_2a_3x3, y1)
            y2, b_2b = unit(i3d.Conv3d_2b_1x1, p1)
            y3, b_2c = unit(i3d.Conv3d_2c_3x3, y2)                            # cross56
            x, b_p2 = pool(i3d.MaxPool3d_3a_3x3, y3)
            chain = []
            for name in ("Mixed_3b", "Mixed_3c", "MaxPool3d_4a_3x3", "Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e", "Mixed_4f"):
                mod = getattr(i3d, name)
                x, bw = pool(mod, x) if name.startswith("MaxPool") else incep(mod, x)
                chain.append(bw)
            N2 = x.shape[0]
            drop1 = E.dropout_scale(N2, 832, dev)
            drop2 = E.dropout_scale(N2, 128, dev)
            ctx_d = _Ctx((True, False))
            xe = E.ChannelScaleFn.forward(ctx_d, x, drop1)
            pc = model.primary_caps
            ctx_pc = _Ctx((True, True, True, True, True, False))
            caps5 = E.PrimaryCapsFn.forward(ctx_pc, xe, pc.pose.weight, pc.pose.bias, pc.a.weight, pc.a.bias, pc)
            caps = caps5[:, 0]
            cc = model.conv_caps
            ctx_r = _Ctx((True, True, True, True))
            rout = E.EMRoutingFn.forward(ctx_r, caps, cc.weights, cc.beta_u, cc.beta_a)
            ctx_a = _Ctx((True,))
            act = E.ClassActFn.forward(ctx_a, rout)
            feat = rout[..., C * 16:].reshape(N2, -1, C)
            cls2 = torch.cat([action, action]).to(dev)
            lab2 = torch.cat([labels_dev, labels_dev])
            lab = torch.nn.functional.one_hot(cls2.long().view(-1), C).float()
            # unlabeled clips: all-ones before thresh_epoch, arg-max one-hot afterwards (capsules_ucf101.py:462-470);
            # the switch is the device scalar hyper[7], so the captured graph does not depend on the epoch
            pseudo = hyper[7]
            unl = pseudo * torch.nn.functional.one_hot(torch.argmax(act, dim=1), C).float() + (1.0 - pseudo)
            sel = (lab2.view(-1, 1) == 0).float()
            mask = (sel * unl + (1.0 - sel) * lab).contiguous()
            ctx_h = _Ctx((True, False))
            x0 = E.CapsHeadFn.forward(ctx_h, rout, mask)
            ctx_dec = _Ctx((True, True, True, True, False, False) + (True,) * 16)
            dparams = []
            for n in E.DecoderFn.ORDER:
                m = getattr(model, n)
                dparams += [m.weight, m.bias]
            logits = E.DecoderFn.forward(ctx_dec, x0, xe, y3, y1, drop2, model, *dparams)
        finally:
            E.STATE.bn_groups = 1

        # ---------------- losses + their gradients (device only) ----------------
        out, flp = logits[:P], logits[P:]
        V = out.numel() // P
        dlogits = torch.zeros_like(logits)
        dact = torch.zeros_like(act)
        seg = ops.u8_to_f32(seg.contiguous()) if seg.dtype == torch.uint8 else seg.contiguous().float()
        sums = torch.empty(4, dtype=torch.float64, device=dev)
        l_seg = torch.zeros(2, dtype=torch.float32, device=dev)
        l_cls = torch.zeros(2, dtype=torch.float32, device=dev)
        if n_lab > 0:
            ops.seg_loss_fwd(out, seg, lab_idx, n_lab, V, sums, l_seg)
            ops.seg_loss_bwd(out, seg, lab_idx, n_lab, V, sums, a.wt_loc, a.wt_loc, dlogits)
            ops.spread_loss(act, action.to(dev).float().contiguous().view(-1), lab_idx, n_lab, act.shape[1], 0.2, l_cls,
                            a.wt_cls, dact)
        m_clk = m_anti = m_gv = None
        if a.bv:
            m_clk = torch.empty((P, 8, H, W), dtype=torch.float32, device=dev)
            m_anti = torch.empty((P, 8, H, W), dtype=torch.float32, device=dev)
            mm = torch.empty((P, 2), dtype=torch.float32, device=dev)
            # clock: pred = output, flip_pred = flipT(flipW(flip_op)) ; anticlock: pred = flipT(output), flip_pred = flipW(flip_op)
            ops.bv_mask(out, flp, m_clk, mm, P, H, W, a.n_frames, a.predict_maps, 0, 1, 1)
            ops.bv_mask(out, flp, m_anti, mm, P, H, W, a.n_frames, a.predict_maps, 1, 0, 1)
        if a.gv:
            m_gv = torch.empty((P, 8, H, W), dtype=torch.float32, device=dev)
            mm2 = torch.empty((P, 2), dtype=torch.float32, device=dev)
            ops.gv_mask(out, m_gv, mm2, P, H, W, a.lower_thresh, a.upper_thresh)
        acc = torch.empty(4, dtype=torch.float64, device=dev)
        l_cons = torch.empty(4, dtype=torch.float32, device=dev)
        mode = (1 if a.bv else 0) | (2 if a.gv else 0)
        ops.cons_reduce(out, flp, m_clk, m_anti, m_gv, acc, P, H, W, 1, 1)
        ops.cons_finish(acc, l_cons, P, H, W, mode, 0.0, 0.0, 0.0, dev_scalars=hyper[1:4])
        ops.cons_grad(out, flp, m_clk, m_anti, m_gv, dlogits[:P], dlogits[P:], P, H, W, 1, 1, 0.0, 0.0, 0.0,
                      dev_scalars=hyper[4:7])

        # ---------------- backward (explicit reverse walk) ----------------
        try:
            gd = E.DecoderFn.backward(ctx_dec, dlogits)
            dx0, dxe_dec, dy3_dec, dy1_dec = gd[0], gd[1], gd[2], gd[3]
            drout = E.CapsHeadFn.backward(ctx_h, dx0)[0]
            drout.add_(E.ClassActFn.backward(ctx_a, dact))                   # fp32 (2P,20,20,408): tiny
            dcaps = E.EMRoutingFn.backward(ctx_r, drout)[0]
            dxe = E.PrimaryCapsFn.backward(ctx_pc, dcaps.view(caps5.shape))[0]
            ops.add(View(dxe), View(dxe_dec), View(dxe))                      # two consumers of the encoder output
            if self.world > 1 and not split:
                # everything after the encoder (84 % of the parameters, incl. the 138 MB PrimaryCaps weight) is final now:
                # all-reduce it on the side stream underneath the encoder backward
                self.buckets.allreduce(1)
            g = E.ChannelScaleFn.backward(ctx_d, dxe)[0]
            for bw in reversed(chain):
                g = bw(g)
            g = b_p2(g)
            ops.add(View(g), View(dy3_dec), View(g))                          # cross56 skip
            g = b_2c(g)
            g = b_2b(g)
            g = b_p1(g)
            ops.add(View(g), View(dy1_dec), View(g))                          # cross112 skip
            b_stem(g)
        finally:
            E.STATE.direct_grads = False
            E.STATE.zero_arena = None
            E.STATE.defer_bn_counters = False
            ops.PACKS = None
        if self.bn_counters:
            torch._foreach_add_(self.bn_counters, 2)      # every BatchNorm saw two training forwards (main_ucf101.py:85-86)
        if not split:
            if self.world > 1:
                self.buckets.allreduce(0)
                self.buckets.join()
            self._optimizer()
        loc = l_seg[0] + l_seg[1]
        total = a.wt_loc * loc + a.wt_cls * l_cls[0] + a.wt_cons * l_cons[0]
        return dict(total=total, loc=loc, bce=l_seg[0], dice=l_seg[1], cls=l_cls[0], cons=l_cons[0], l2=l_cons[1],
                    output=out, flip_op=flp, pred_action=act[:P], feat=feat[:P])


class _Ctx:
    """Minimal stand-in for the autograd context so the engine functions' forward / backward halves can be called
    directly by the hand-scheduled step."""

    def __init__(self, needs_input_grad):
        self.needs_input_grad = tuple(needs_input_grad)
        self.saved_tensors = ()

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors
