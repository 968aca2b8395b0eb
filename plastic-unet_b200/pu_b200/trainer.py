"""The batched train step of the hot path: forward + BCE + backward + (DP all-reduce) + Adam + plastic-trace
carry, captured once into a CUDA graph and replayed (no tracing compiler; reference loop: train.py:91-112).

All parameters live in ONE flat arena (and their gradients in another), so the optimizer is a single
pu_adam_step launch and the data-parallel gradient exchange is a single all-reduce.
"""
import os

import torch
import torch.distributed as dist

from . import _lib, ops

_os_environ_get = os.environ.get


class TrainStep:
    """step(x, target) -> device scalar loss.  x: [B, C, H, W] float32 (device or pinned host), target: [B, nbf, nbf].

    The trace `hebb` is carried (detached) from step to step like train.py:99 and reset by reset_trace()
    (train.py:88 resets it every epoch)."""

    def __init__(self, net, batch, in_hw, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, use_graph=True, dp_group=None, warmup=3):
        self.net = net
        self.dev = next(net.parameters()).device
        self.batch = batch
        self.world = dist.get_world_size(dp_group) if (dist.is_initialized() and dp_group is not None) else 1
        self.dp_group = dp_group if self.world > 1 else None
        self.betas, self.eps = betas, eps
        params = [p for p in net.parameters()]
        total = sum(p.numel() for p in params)
        # Data parallel, opt-in (PU_DP_BUCKETS=1): two gradient buckets.  The 1.06 MB exchange is latency-bound, so two all-reduces
        # cost more than the overlap hides (4 GPUs: +19 us per step); kept for models with more parameters.  The parameters of the shallow encoder levels (inc, down1 .. down{L-1}) get their
        # gradients LAST in the backward pass; everything else (head, decoder, deep encoder levels: ~90 % of the bytes) is
        # complete when the gradient w.r.t. the input of level L exists, and is all-reduced on a communication stream while
        # the shallow levels still run.  The arena is ordered [early bucket | late bucket].
        self._late = set()
        self._bucket_level = 0
        if self.dp_group is not None and hasattr(net, "_bucket_hook") and type(net).__name__ == "UNetp" and getattr(net, "depth", 0) >= 3 \
                and os.environ.get("PU_DP_BUCKETS", "0") == "1":  # measured on 4 B200s: 1.238 ms with, 1.220 ms without -> opt-in
            self._bucket_level = net.depth - 1
            late_prefixes = ("inc.",) + tuple("down%d." % j for j in range(1, self._bucket_level))
            self._late = {id(p) for n_, p in net.named_parameters() if n_.startswith(late_prefixes)}
            params = [p for p in params if id(p) not in self._late] + [p for p in params if id(p) in self._late]
        # ---- flat arenas (32-byte aligned slots: the conv epilogues read biases with 256-bit loads)
        self.offsets = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + 7) // 8 * 8
        self.n_flat = off
        self.split = next((o for p, o in zip(params, self.offsets) if id(p) in self._late), off)  # first float of the late bucket
        self._comm = torch.cuda.Stream() if self._late else None
        self._early_done = False
        self._grad_params = None  # learnt by the first eager step: which parameters receive a gradient at all
        self.flat_p = torch.zeros(off, device=self.dev)
        self.flat_g = torch.zeros(off, device=self.dev)
        # Data parallel: the gradient exchange fused with Adam over NVLink peer memory (pu_adam_allreduce_step): the gradient arena
        # and a flag buffer live in symmetric memory; falls back to ncclAllReduce + pu_adam_step if that is unavailable.  The
        # plastic-trace delta rides along (no NCCL kernel inside the step at all).  Default at 2, 4 and 8 GPUs (measured 5 / 7 / 13 us
        # per step faster than NCCL, replicas bit-identical); PU_DP_FUSED=0 restores NCCL.
        self._fused = None
        want_fused = os.environ.get("PU_DP_FUSED", "1") == "1"
        if self.dp_group is not None and self.world in (2, 4, 8) and not self._late and want_fused:
            try:
                import torch.distributed._symmetric_memory as symm
                nblk = int(_lib.load().pu_adam_allreduce_blocks())
                # the arena is followed by the plastic-trace delta of the step (N*N + N floats, pu_trace_delta): summed over the
                # ranks by the same kernel, so the trace all-reduce needs no NCCL kernel in the middle of the backward pass
                N = net.nbf
                self._n_trace = (N * N + N + 3) // 4 * 4
                g_sym = symm.empty(off + self._n_trace, dtype=torch.float32, device=self.dev)
                g_sym.zero_()
                f_sym = symm.empty(2 * nblk * self.world, dtype=torch.int32, device=self.dev)
                f_sym.zero_()
                hg = symm.rendezvous(g_sym, self.dp_group)
                hf = symm.rendezvous(f_sym, self.dp_group)
                torch.cuda.synchronize()
                dist.barrier(self.dp_group)  # every rank's flags are zero before anyone signals
                self._fused = (hg, hf, f_sym, int(hg.rank))
                self._arena = g_sym
                self.flat_g = g_sym[:off]
                self._delta_sum = torch.zeros(self._n_trace, device=self.dev)
                self._eta_snap = torch.zeros(1, device=self.dev)
                self._trace_job = None
            except Exception as e:  # no peer access / symmetric memory on this system
                import warnings
                warnings.warn("pu_b200.TrainStep: fused peer-memory gradient exchange unavailable (%s: %s); using NCCL all-reduce"
                              % (type(e).__name__, str(e)[:200]))
                self._fused = None
        self.m = torch.zeros(off, device=self.dev)
        self.v = torch.zeros(off, device=self.dev)
        self.params = params
        for p, o in zip(params, self.offsets):
            self.flat_p[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + p.numel()].view_as(p)
            p.grad = None
        # gradient sink (ops.GRAD_SINK): backward ops whose kernels can accumulate write dw / db straight into their arena slots
        # (zeroed once per step): no memset launches per op, and those parameters need no gather.  PU_GRAD_SINK=0 disables.
        self._sink = None
        if os.environ.get("PU_GRAD_SINK", "0") == "1":
            slot = {id(p): (o, p.numel()) for p, o in zip(params, self.offsets)}
            self._sink = {}
            for m in net.modules():
                w = getattr(m, "weight", None)
                if isinstance(w, torch.nn.Parameter) and w.dim() == 4 and id(w) in slot:
                    b = getattr(m, "bias", None)
                    has_b = isinstance(b, torch.nn.Parameter) and id(b) in slot
                    self._sink[w.data_ptr()] = (self.flat_g, slot[id(w)][0], slot[id(b)][0] if has_b else -1, slot[id(b)][1] if has_b else 0)
        # (pointer, offset, size) table of the gradient tensors, rebuilt by every eager/captured step body
        self.table_host = torch.zeros((len(params), 3), dtype=torch.int64).pin_memory()
        self.table_dev = torch.zeros((len(params), 3), dtype=torch.int64, device=self.dev)
        self.table2_host = torch.zeros((len(params), 3), dtype=torch.int64).pin_memory()  # the early bucket's table
        self.table2_dev = torch.zeros((len(params), 3), dtype=torch.int64, device=self.dev)
        self.n_params = total
        self.step_count = torch.zeros(1, device=self.dev)
        self.lr = torch.full((1,), float(lr), device=self.dev)
        C = net.n_channels
        self.x = torch.zeros((batch, C, in_hw, in_hw), device=self.dev)
        self.target = torch.zeros((batch, net.nbf, net.nbf), device=self.dev)
        self.hebb = net.initialZeroHebb()
        self.loss = torch.zeros(1, device=self.dev)
        self._unit = torch.ones(1, device=self.dev)  # seed of loss.backward() in the fused-head form
        self._head_fused = os.environ.get("PU_HEAD_FUSED", "1") == "1"
        self.graph = None
        self._table_early = False
        self.kernels_per_step = None
        self.use_graph = use_graph
        self._warm = warmup
        self._pack_list = None  # learnt by the first eager step: [(weight, transpose, math, C0)]
        self._pack_stream = None
        # PU_ADAM_TABLE=1 (single GPU, no gradient sink): pu_adam_table_step reads the gradient tensors directly — one launch for gather +
        # step counter + Adam.  Measured SLOWER (1.017 vs 1.002 ms per step: one block column per tensor serialises the large
        # tensors and wastes blocks on the 8-element biases), so it stays opt-in
        self._adam_table = self.dp_group is None and self._sink is None and os.environ.get("PU_ADAM_TABLE", "0") == "1"
        self._aux_stream = None   # uploads the gather's pointer table beside the step's main chain (captured steps)
        self._aux_pending = False
        import os as _os
        # weight gradients run on a side stream and overlap the dgrad chain (measured 1.56 -> 1.38 ms/step); PU_WGRAD_SIDE=0 disables
        nside = int(_os.environ.get("PU_WGRAD_SIDE", "4"))  # number of side streams (round-robin); 0 = none; measured 1: 1.305, 2: 1.191, 4: 1.171, 8: 1.174 ms
        self.wgrad_side = [torch.cuda.Stream() for _ in range(nside)] if nside > 0 else None

    # -------------------------------------------------------------------------------------------
    def _prepack(self):
        """Pack the weight images that do not fit in shared memory (the deep 64-channel layers) on a side stream at the
        start of the step, for the forward and for the dgrad convs alike: off the critical path (10 launches, ~35 us)."""
        if not self._pack_list:
            ops.PACK_CACHE = None
            return
        if self._pack_stream is None:
            self._pack_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        self._pack_stream.wait_stream(main)
        cache = {}
        with torch.cuda.stream(self._pack_stream):
            for (w, transpose, math, C0) in self._pack_list:
                wp = ops._pack_w_now(w, transpose, math, C0)
                ev = torch.cuda.Event()
                ev.record(self._pack_stream)
                wp.record_stream(main)
                cache[(w.data_ptr(), transpose, math, C0)] = (wp, ev)
        ops.PACK_CACHE = cache

    def _step_body(self):
        st = torch.cuda.current_stream().cuda_stream
        for p in self.params:
            p.grad = None  # autograd then hands over its gradient tensors instead of launching one add per parameter
        # Under capture the pointer table of the gradient gather is uploaded at the START of the step (a memcpy node reads the
        # pinned host table at replay time, and the table is final when the capture ends): 3 us off the step's tail.
        # The copy runs on its own stream, joined only by the gather at the end of the step: as the first node of the main chain
        # the copy-engine -> SM hand-off put ~12 us in front of the first kernel (profiles/r2_step_timeline_side4.txt).
        self._table_early = torch.cuda.is_current_stream_capturing()
        if self._table_early:
            if self._aux_stream is None:
                self._aux_stream = torch.cuda.Stream()
            self._aux_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._aux_stream):
                self.table_dev.copy_(self.table_host, non_blocking=True)
            self._aux_pending = True
        if self._sink is not None:
            self.flat_g.zero_()  # the gradient kernels accumulate into the arena
        # (without the sink the gather overwrites every slot that has a gradient; the others and the alignment gaps stay zero
        # from the construction on — nothing else writes the arena, the all-reduce adds zeros to them)
        if self._pack_list is None:
            ops.PACK_LOG, ops.PACK_CACHE = [], None  # first eager step: learn which weights the step packs
        else:
            self._prepack()
        try:
            self._fwd_bwd(st)
        finally:
            if self._pack_list is None:
                seen, self._pack_list = set(), []
                for (w, transpose, math, C0) in ops.PACK_LOG:
                    k = (w.data_ptr(), transpose, math, C0)
                    if k not in seen:
                        seen.add(k)
                        self._pack_list.append((w, transpose, math, C0))
            ops.PACK_LOG = ops.PACK_CACHE = None
            if self._pack_stream is not None:
                torch.cuda.current_stream().wait_stream(self._pack_stream)  # join (needed under graph capture)

    def _fill_table(self, which, table_host):
        """(gradient pointer, arena offset, element count) rows of the parameters selected by `which(p)` -> their number."""
        n = 0
        base = self.flat_g.data_ptr()
        for p, o in zip(self.params, self.offsets):
            if p.grad is not None and which(p):
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if g.data_ptr() == base + 4 * o:
                    continue  # written in place through the gradient sink
                table_host[n, 0], table_host[n, 1], table_host[n, 2] = g.data_ptr(), o, g.numel()
                n += 1
        return n

    def _gather(self, which, table_host, table_dev, inc=False):
        """One-launch gather of the gradients of the parameters selected by `which(p)` into their flat_g slots.  inc: the same
        launch increments the optimizer's step counter; returns whether it did."""
        n = self._fill_table(which, table_host)
        if n:
            if not (self._table_early and table_dev is self.table_dev):
                table_dev.copy_(table_host, non_blocking=torch.cuda.is_current_stream_capturing())
            st = torch.cuda.current_stream().cuda_stream
            if inc:
                _lib.call("pu_gather_flat_inc", table_dev.data_ptr(), n, self.flat_g.data_ptr(), self.step_count.data_ptr(), st)
                return True
            _lib.call("pu_gather_flat", table_dev.data_ptr(), n, self.flat_g.data_ptr(), st)
        return False

    def _early_reduce(self, _grad):
        """Backward hook (modules.UNetp.forward): the early bucket is complete -> gather + all-reduce it on the communication
        stream, concurrently with the rest of the backward pass.  Falls back to the single all-reduce if a gradient of the
        bucket has not been handed over yet."""
        early = [p for p in self.params if id(p) not in self._late and id(p) in self._grad_params]
        if any(p.grad is None for p in early):
            return None
        main = torch.cuda.current_stream()
        self._comm.wait_stream(main)
        for sd in (self.wgrad_side or []):
            self._comm.wait_stream(sd)  # their weight-gradient kernels write the tensors gathered below
        with torch.cuda.stream(self._comm):
            self._gather(lambda p: id(p) not in self._late, self.table2_host, self.table2_dev)
            dist.all_reduce(self.flat_g[:self.split], op=dist.ReduceOp.SUM, group=self.dp_group)
        self._early_done = True
        return None

    def _fwd_bwd(self, st):
        self._early_done = False
        self.net._bucket_hook = self._early_reduce if (self._late and self._grad_params is not None) else None
        if self._late:
            self.net._bucket_level = self._bucket_level
        # the trace update (data parallel: the trace-delta all-reduce + epilogue) runs on net.dp_side, off the critical path; only
        # this step body defers it (it joins dp_side before Adam) — any other caller of net.forward gets the trace on its own stream
        self.net.dp_defer = True
        if self._fused is not None:
            self._trace_job = None
            self.net.dp_external = self._take_trace_delta
        # TF32 mode: the head, the BCE loss and the backward of both are ONE kernel inside the forward pass (modules._plastic);
        # PU_HEAD_FUSED=0 keeps the separate head GEMMs + pu_bce_fwd_bwd (always used by the strict-fp32 and mixed modes)
        fuse = self._head_fused and self.net.nbf <= 128 and getattr(self.net, "conv_math", "") == "tf32"
        self.net._head_bce = self.target if fuse else None
        if fuse:
            # Weff = w + alpha*hebb depends only on what is known at the start of the step: computed beside the forward pass, so
            # that the fused head kernel stages 64 KB by asynchronous copies instead of building it from three tensors
            main = torch.cuda.current_stream()
            if self._aux_stream is None:
                self._aux_stream = torch.cuda.Stream()
            self._aux_stream.wait_stream(main)
            with torch.cuda.stream(self._aux_stream):
                weff = ops.head_weff(self.net.w.detach(), self.net.alpha.detach(), self.hebb)
                ev = torch.cuda.Event()
                ev.record(self._aux_stream)
            weff.record_stream(main)
            self.net._head_weff = (weff, ev)
            self._aux_pending = True  # joined before the gradient gather
        try:
            out, hebb_new = self.net(self.x, self.hebb)
        finally:
            self.net.dp_defer = False
            self.net.dp_external = None
            self.net._head_bce = None
            self.net._head_weff = None
        fused_loss, self.net._head_loss = self.net._head_loss, None
        if fused_loss is not None:
            self.loss = fused_loss.detach()  # under capture: a tensor of the graph's pool, rewritten by every replay
            root, seed = fused_loss, self._unit
        else:
            gS = torch.empty_like(out)
            n = out.numel()
            _lib.call("pu_bce_fwd_bwd", out.data_ptr(), self.target.data_ptr(), self.loss.data_ptr(), gS.data_ptr(), n, st)
            root, seed = out, gS
        check = getattr(self, "_check_grads", False) and self.wgrad_side is not None
        if self.wgrad_side is not None:
            ops.WGRAD_SIDE_STREAMS = self.wgrad_side
        if check:
            ops.SIDE_OUTPUTS = set()
        ops.GRAD_SINK = self._sink
        ops.UNIT_GRAD = self._unit
        try:
            root.backward(seed)
        finally:
            ops.UNIT_GRAD = None
            ops.GRAD_SINK = None
            ops.WGRAD_SIDE_STREAMS = None
            side_ptrs, ops.SIDE_OUTPUTS = ops.SIDE_OUTPUTS, None
        if check:
            # The side-stream gradient kernels are ordered against the main stream only by the join below.  That is safe
            # iff autograd handed the very tensors those kernels write over as .grad (no copy / accumulation kernel on
            # the main stream in between): assert it once, on the first eager warm-up step.
            got = {p.grad.data_ptr() for p in self.params if p.grad is not None}
            if not side_ptrs <= got:
                raise RuntimeError("TrainStep: %d side-stream parameter gradients were copied or accumulated by autograd "
                                   "instead of being handed over (grad hooks / retain_graph?); set PU_WGRAD_SIDE=0"
                                   % len(side_ptrs - got))
        if self.wgrad_side is not None:
            for sd in self.wgrad_side:
                torch.cuda.current_stream().wait_stream(sd)  # join the side-stream parameter gradients
        # gather every parameter gradient into the flat arena with one launch (parameters without a gradient, e.g.
        # eta in the reference loop, keep their zeroed slot)
        self.net._bucket_hook = None
        if self._grad_params is None:
            self._grad_params = {id(p) for p in self.params if p.grad is not None}
        if self._aux_pending:
            torch.cuda.current_stream().wait_stream(self._aux_stream)  # the table upload started at the top of the step
            self._aux_pending = False
        if self._early_done:
            self._gather(lambda p: id(p) in self._late, self.table_host, self.table_dev)
            dist.all_reduce(self.flat_g[self.split:], op=dist.ReduceOp.SUM, group=self.dp_group)
            torch.cuda.current_stream().wait_stream(self._comm)  # the early bucket's all-reduce
        elif self._adam_table:
            # single GPU: Adam reads the gradient tensors themselves (one launch instead of gather + step counter + Adam)
            n = self._fill_table(lambda p: True, self.table_host)
            if not self._table_early:
                self.table_dev.copy_(self.table_host, non_blocking=torch.cuda.is_current_stream_capturing())
            if getattr(self.net, "dp_side", None) is not None:
                torch.cuda.current_stream().wait_stream(self.net.dp_side)  # the trace epilogue reads eta from the arena
            if n:
                _lib.call("pu_adam_table_step", self.table_dev.data_ptr(), n, self.flat_p.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                          self.step_count.data_ptr(), self.lr.data_ptr(), self.betas[0], self.betas[1], self.eps, 1.0, st)
            self.hebb.copy_(hebb_new.detach())
            return
        else:
            counted = False
            if self._fused is None:
                # The tail of the step is a chain of small launches: the carry of the new trace into self.hebb leaves it (side stream,
                # forked here — every reader of this step's trace, the head's parameter gradients and the Weff computation, has been
                # joined above — and joined at the end), and the gather launch also increments the optimizer's step counter.
                if getattr(self.net, "dp_side", None) is not None:
                    torch.cuda.current_stream().wait_stream(self.net.dp_side)  # the trace update (it reads eta from the arena)
                if self._aux_stream is None:
                    self._aux_stream = torch.cuda.Stream()
                main = torch.cuda.current_stream()
                self._aux_stream.wait_stream(main)
                with torch.cuda.stream(self._aux_stream):
                    self.hebb.copy_(hebb_new.detach())
                hebb_new.record_stream(self._aux_stream)
                counted = self._gather(lambda p: True, self.table_host, self.table_dev, inc=True)
                if self.dp_group is not None:
                    dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.dp_group)
                _lib.call("pu_adam_step_counted" if counted else "pu_adam_step", self.flat_p.data_ptr(), self.flat_g.data_ptr(),
                          self.m.data_ptr(), self.v.data_ptr(), self.step_count.data_ptr(), self.lr.data_ptr(), self.betas[0], self.betas[1],
                          self.eps, 1.0 / self.world, self.n_flat, st)
                main.wait_stream(self._aux_stream)
                return
            self._gather(lambda p: True, self.table_host, self.table_dev)
        if getattr(self.net, "dp_side", None) is not None:
            # join the deferred trace all-reduce + epilogue BEFORE the optimizer rewrites the arena (the epilogue reads eta)
            torch.cuda.current_stream().wait_stream(self.net.dp_side)
        if self._fused is not None:
            hg, hf, _flags, grank = self._fused
            job = self._trace_job
            if job is not None:
                self._eta_snap.copy_(self.net.eta.detach())  # the trace epilogue uses the step's eta, not the optimizer's result
            _lib.call("pu_adam_allreduce_step", self.flat_p.data_ptr(), int(hg.buffer_ptrs_dev), int(hf.buffer_ptrs_dev), grank, self.world,
                      self.m.data_ptr(), self.v.data_ptr(), self.step_count.data_ptr(), self.lr.data_ptr(), self.betas[0], self.betas[1],
                      self.eps, 1.0 / self.world, self.n_flat, self._delta_sum.data_ptr() if job is not None else None,
                      self._n_trace if job is not None else 0, st)
            if job is not None:
                # identical summed delta on every rank -> identical trace; in place (element-wise)
                k_global, rule = job
                _lib.call("pu_trace_apply", self.hebb.data_ptr(), self._delta_sum.data_ptr(), k_global, self._eta_snap.data_ptr(), rule,
                          self.hebb.data_ptr(), self.net.nbf, st)
                self._trace_job = None
                return
        else:
            _lib.call("pu_adam_step", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                      self.step_count.data_ptr(), self.lr.data_ptr(), self.betas[0], self.betas[1], self.eps,
                      1.0 / self.world, self.n_flat, st)
        self.hebb.copy_(hebb_new.detach())

    def _take_trace_delta(self, delta_q, k_global, rule):
        """modules._plastic hands the step's local trace delta over (fused exchange): it goes behind the gradients in the arena."""
        n = delta_q.numel()
        self._arena[self.n_flat:self.n_flat + n].copy_(delta_q.reshape(-1))
        self._trace_job = (int(k_global), int(rule))

    def _state_tensors(self):
        """Every tensor a step mutates: the optimizer state, the trace and the BatchNorm buffers."""
        return [self.flat_p, self.m, self.v, self.step_count, self.hebb] + [b for b in self.net.buffers()]

    def capture(self):
        """Warm up on a side stream, then capture one step into a CUDA graph.  The warm-up runs real steps (on the
        zero-filled static buffers), so the optimizer state, the trace and the BN buffers are snapshotted first and
        restored afterwards: capture() leaves the model exactly as it found it (step count 0, trace untouched)."""
        saved = [t.detach().clone() for t in self._state_tensors()]

        def restore():
            for t, s0 in zip(self._state_tensors(), saved):
                t.copy_(s0)

        # the step's main stream gets the HIGHEST priority: when weight-gradient kernels on the (default-priority) side
        # streams share the GPU with the persistent one-CTA-per-SM conv kernels of the dgrad chain, the block scheduler
        # then hands every SM that frees up to the critical path first
        prio = int(_os_environ_get("PU_MAIN_PRIORITY", "-1"))
        s = torch.cuda.Stream(priority=prio)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for i in range(self._warm):
                self._check_grads = (i == 0)
                self._step_body()
        self._check_grads = False
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if not self.use_graph:
            before = _lib.launch_count()
            self._step_body()
            self.kernels_per_step = _lib.launch_count() - before
            torch.cuda.synchronize()
            restore()
            torch.cuda.synchronize()
            return self
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph, stream=s):
            self._step_body()
        self.kernels_per_step = _lib.launch_count() - before
        restore()
        torch.cuda.synchronize()
        return self

    def reset_trace(self):
        self.hebb.zero_()

    def set_lr(self, lr):
        self.lr.fill_(float(lr))

    def _load(self, x, target):
        """Copy a batch (host or device tensors; either may be None) into the step's static buffers."""
        if (x is not None and target is not None and x.is_cuda and target.is_cuda and x.dtype == torch.float32
                and target.dtype == torch.float32 and x.is_contiguous() and target.is_contiguous() and x.numel() == self.x.numel()
                and target.numel() == self.target.numel() and x.numel() % 4 == 0 and target.numel() % 4 == 0
                and x.device == self.x.device and target.device == self.x.device
                and (x.data_ptr() | target.data_ptr() | self.x.data_ptr() | self.target.data_ptr()) % 16 == 0):
            # device-resident batch: both copies in one launch (two memcpy calls cost two launch gaps in front of every replay)
            _lib.call("pu_copy2", x.data_ptr(), self.x.data_ptr(), x.numel(), target.data_ptr(), self.target.data_ptr(), target.numel(),
                      torch.cuda.current_stream().cuda_stream)
        else:
            if x is not None:
                self.x.copy_(x, non_blocking=True)
            if target is not None:
                self.target.copy_(target, non_blocking=True)

    def step(self, x=None, target=None):
        """One optimisation step.  Host (pinned) or device tensors are copied into the static buffers first.  Returns the
        step's loss tensor (device, [1]); use the returned tensor or read `self.loss` AFTER the call — in the TF32 mode the
        fused head kernel writes the loss into a tensor of its own, so `self.loss` is rebound by capture() and by eager steps."""
        self._load(x, target)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()
        return self.loss

    def step_indices(self, dataset, idx):
        """One optimisation step on samples `idx` ([B] int64; host or device) of a pu_b200.data.DeviceDataset: the batch is
        gathered and zero-padded on the device (two launches), only B indices cross PCIe."""
        if getattr(self, "_idx_dev", None) is None:
            self._idx_dev = torch.zeros(self.batch, dtype=torch.int64, device=self.dev)
        self._idx_dev.copy_(idx, non_blocking=True)
        dataset.gather(self._idx_dev, self.x, self.target)
        return self.step()

    # -------------------------------------------------------------------------------------------
    def prefetch(self, x, target):
        """Start the host->device copy of the NEXT batch on a copy stream; it overlaps the step that is running.
        Pair with step_prefetched().  (Pinned host tensors; the staging buffers are reused once consumed.)"""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage_x = torch.empty_like(self.x)
            self._stage_t = torch.empty_like(self.target)
            self._stage_ready = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free)  # the previous staged batch has been moved into the step's buffers
            self._stage_x.copy_(x, non_blocking=True)
            self._stage_t.copy_(target, non_blocking=True)
            self._stage_ready.record(self._copy_stream)

    def step_prefetched(self):
        """One optimisation step on the batch handed to prefetch()."""
        main = torch.cuda.current_stream()
        main.wait_event(self._stage_ready)
        self._load(self._stage_x, self._stage_t)  # staged batch -> static buffers (one launch)
        self._stage_free.record(main)              # the staging buffers may be refilled while the step runs
        return self.step()


class InferStep:
    """Batched forward-only step (eval.py:81-90 / infer.py:42-47 semantics: trace zero, hebb' discarded)."""

    def __init__(self, net, batch, in_hw, use_graph=True):
        self.net = net  # run in eval mode inside capture()/step(); the caller's training flag is restored afterwards
        self.dev = next(net.parameters()).device
        self.x = torch.zeros((batch, net.n_channels, in_hw, in_hw), device=self.dev)
        self.hebb = net.initialZeroHebb()
        self.out = None
        self.graph = None
        self.use_graph = use_graph
        self.kernels_per_step = None

    class _eval_mode:
        def __init__(self, net):
            self.net = net

        def __enter__(self):
            self.was = self.net.training
            self.net.eval()

        def __exit__(self, *exc):
            self.net.train(self.was)

    def capture(self):
        with torch.no_grad(), InferStep._eval_mode(self.net):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    self.out, _ = self.net(self.x, self.hebb)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            before = _lib.launch_count()
            if self.use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self.out, _ = self.net(self.x, self.hebb)
            else:
                self.out, _ = self.net(self.x, self.hebb)
            self.kernels_per_step = _lib.launch_count() - before
        return self

    def step(self, x=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad(), InferStep._eval_mode(self.net):
                self.out, _ = self.net(self.x, self.hebb)
        return self.out

    def predict_rle(self, x=None, mask_threshold=0.5, want_mask=False):
        """infer.predict for a batch (infer.py:73-99): forward, `mask > mask_threshold`, column-major RLE strings — masks and
        run lists are produced on the device (pu_b200.infer_tail), only the runs cross PCIe."""
        from . import infer_tail
        out = self.step(x)
        out = out.view(-1, self.net.nbf, self.net.nbf)
        return infer_tail.rle_encode_batch(out, mask_threshold, want_mask=want_mask)
