"""NVLink-fused expert-parallel engine: dispatch+GEMM1 and GEMM2+combine with no separate all-to-all.

This is the B200-native replacement of the reference's ``all_to_all -> experts -> all_to_all`` sequence
(tutel/impls/moe_layer.py:349-351) and of its chunked NCCL overlap scheduler (tutel/impls/overlap.py,
tutel/custom/custom_kernel.cpp:520-654).  Per forward pass and rank (W ranks, El local experts, capacity C):

  comm stream   encode kernel: gathers this rank's tokens slot by slot and *stores them straight into the
                destination expert GPU's receive buffer* ``X_recv[El, W(src), C, M]`` over NVLink, publishing an
                epoch flag per row chunk with ``st.release.sys``                       (csrc/moe_kernels.cu)
  main stream   GEMM1 (tcgen05): its TMA producer ``ld.acquire.sys``-polls the flags of exactly the rows of the
                tile it is about to load, so tiles are multiplied as they arrive - own-rank rows first.
                GEMM2 (tcgen05): the epilogue stores every output tile *directly into the source GPU's* combine
                buffer ``Y_comb[E, C, Mout]`` and bumps a ``red.release.sys`` counter per expert.
                decode kernel: acquires the counters of the experts a token used and sums its k rows.

Backward mirrors this (output gradients are dispatched, dgrad/wgrad GEMMs run as rows arrive, input gradients are
combined), so one training step issues 4 fused transfers and never calls NCCL for tokens.
``a2a_ffn_overlap_degree`` selects the flag granularity (rows per arrival flag = C / d, at least one MMA tile).

Buffers live in the symmetric heap (parallel/p2p.py).  Each layer owns a small ring of buffer sets; a set stays
reserved from forward until its backward finished and sets are re-used least-recently-used first, which - together
with the data dependencies between ranks - guarantees that no peer can overwrite rows that are still being read.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import torch

from ..ops import backend
from ..ops import gemm as G
from ..ops.dispatch import DispatchPlan
from . import p2p

_FLAGS_PER_SEG = 64          # arrival flags per (expert, source) segment
_MIN_RING = 2
_MAX_RING = 6


def _enabled() -> bool:
    return os.environ.get('TUTEL_B200_FUSED', '1') not in ('0', 'off', 'false')


class _BufferSet:
    """One set of symmetric buffers + flag areas (identical offsets on every rank)."""

    def __init__(self, eng: 'FusedEngine', index: int, C: int):
        t, es = eng.t, eng.es
        W, El, E, M, Mo = eng.W, eng.El, eng.E, eng.M, eng.Mo
        tag = '%s/set%d/C%d' % (eng.tag, index, C)
        self.C = C
        self.x_recv = t.alloc(tag + '/x_recv', El * W * C * M * es)
        self.y_comb = t.alloc(tag + '/y_comb', E * C * Mo * es)
        self.dy_recv = t.alloc(tag + '/dy_recv', El * W * C * Mo * es)
        self.dx_comb = t.alloc(tag + '/dx_comb', E * C * M * es)
        self.f_disp = t.ctrl_alloc(tag + '/f_disp', El * W * _FLAGS_PER_SEG * 4)
        self.f_comb = t.ctrl_alloc(tag + '/f_comb', E * 4)
        self.b_disp = t.ctrl_alloc(tag + '/b_disp', El * W * _FLAGS_PER_SEG * 4)
        self.b_comb = t.ctrl_alloc(tag + '/b_comb', E * 4)
        self.epoch = 0                # published by the encode kernels (one writer per flag)
        self.comb_total = [0, 0]      # cumulative tile counts expected on f_comb / b_comb
        self.busy = False
        self.last_used = -1
        self.tables: Dict[Any, torch.Tensor] = {}

    @staticmethod
    def bytes_needed(eng: 'FusedEngine', C: int) -> int:
        return 2 * (eng.El * eng.W * C * (eng.M + eng.Mo)) * eng.es + 4096


class FusedEngine:
    _serial = 0

    def __init__(self, layer, transport: 'p2p.P2PTransport', dtype: torch.dtype):
        ex = layer.experts
        self.t = transport
        self.W, self.rank = transport.world, transport.rank
        self.El, self.E = layer.num_local_experts, layer.num_global_experts
        if hasattr(ex, 'full_shapes'):       # llama_ffn: W_fc1/W_fc2 [El, M, H], W_fc3 [El, H, M]
            self.M, self.H, self.Mo = layer.model_dim, int(ex.full_shapes['W_fc1'][2]), int(ex.full_shapes['W_fc3'][2])
        else:
            self.M, self.H, self.Mo = layer.model_dim, ex.hidden_size, ex.output_dim
        self.dtype, self.es = dtype, torch.empty((), dtype=dtype).element_size()
        FusedEngine._serial += 1                 # same creation order on every rank -> same names and offsets
        self.tag = 'moe%d' % FusedEngine._serial
        self.sets: Dict[int, List[_BufferSet]] = {}
        self.clock = 0
        self.side = torch.cuda.Stream()
        self.chunk_counters = torch.zeros([self.E * _FLAGS_PER_SEG], dtype=torch.int32, device='cuda')
        self.warned = False

    # ---- buffer ring ------------------------------------------------------------------------------------------
    def acquire(self, C: int, hold: bool) -> Optional[_BufferSet]:
        ring = self.sets.setdefault(C, [])
        free = [s for s in ring if not s.busy]
        if len(ring) < _MIN_RING or not free:
            if len(ring) >= _MAX_RING and not free:
                # every set is still referenced by a forward whose backward has not run (deep micro-batching, weight
                # sharing): its saved rows must not be overwritten, so this call takes the generic all-to-all path
                return None
            if not self.t.can_alloc(_BufferSet.bytes_needed(self, C)):
                return None if not free else self._take(min(free, key=lambda s: s.last_used), hold)
            ring.append(_BufferSet(self, len(ring), C))
            return self._take(ring[-1], hold)
        return self._take(min(free, key=lambda s: s.last_used), hold)

    def _take(self, s: _BufferSet, hold: bool) -> _BufferSet:
        self.clock += 1
        s.last_used, s.busy = self.clock, hold
        return s

    # ---- pointer tables (device int64 arrays, cached per buffer set) ------------------------------------------------
    def _table(self, s: _BufferSet, key, values: List[int]) -> int:
        tab = s.tables.get(key)
        if tab is None:
            tab = torch.tensor(values, dtype=torch.int64, device='cuda')
            s.tables[key] = tab
        return tab.data_ptr()

    def push_tables(self, s: _BufferSet, data_off: int, flag_off: int, width: int):
        """Encode side: expert e's rows go to rank e//El, segment (e%El, my rank)."""
        C, W, El, rank, es = s.C, self.W, self.El, self.rank, self.es
        dst = [self.t.base_ptr(e // El) + data_off + ((e % El) * W + rank) * C * width * es for e in range(self.E)]
        sig = [self.t.base_ptr(e // El) + flag_off + ((e % El) * W + rank) * _FLAGS_PER_SEG * 4 for e in range(self.E)]
        return self._table(s, ('pd', data_off, width), dst), self._table(s, ('ps', flag_off), sig)

    def combine_tables(self, s: _BufferSet, data_off: int, flag_off: int, width: int):
        """GEMM epilogue side: group g = (local expert, source rank) is written into the source rank's buffer."""
        C, W, El, rank, es = s.C, self.W, self.El, self.rank, self.es
        dst, sig = [], []
        for g in range(El * W):
            el, src = divmod(g, W)
            e = rank * El + el
            dst.append(self.t.base_ptr(src) + data_off + e * C * width * es)
            sig.append(self.t.base_ptr(src) + flag_off + e * 4)
        return self._table(s, ('cd', data_off, width), dst), self._table(s, ('cs', flag_off), sig)

    def chunk_rows(self, C: int, d: int) -> int:
        """Rows per arrival flag.  One MMA tile (256 rows) is the finest useful granularity; `d` (the layer's
        a2a_ffn_overlap_degree) can only make it finer than the default C/8, never coarser than C."""
        rows = -(-C // max(d, 8))
        rows = max(256, rows, -(-C // _FLAGS_PER_SEG))
        return (rows + 255) // 256 * 256

    @staticmethod
    def tile_counts(C: int, N: int):
        cg = 2 if C > 128 else 1
        bn = 256 if N > 128 else 128
        return cg, bn, (-(-C // (128 * cg))) * (-(-N // bn))


def engine_for(layer, x: torch.Tensor, crit, d: int):
    """Return a runnable fused engine for this call, or None when the generic path must be used."""
    if not _enabled() or not backend.use_tcgen05(x):
        return None
    ex = layer.experts
    from ..models.experts.ffn import FusedExpertsNetwork
    from ..models.experts.llama_ffn import LlamaFFNNetwork
    if layer.sharded_count != 1 or layer.adaptive_degree != 1 or layer.megablocks_size > 0:
        return None
    if not layer.is_postscore and os.environ.get('TUTEL_B200_FUSED_PRESCORE', '0') != '1':
        # gate-before-experts ("prescore") runs through the same kernels (gates applied in the push, gate gradients from
        # the combined input gradients) but has no multi-GPU equivalence test yet: generic path unless asked for
        return None
    if isinstance(ex, FusedExpertsNetwork):
        if ex._act_kind != 'relu' or ex.skip_expert:
            return None
        if ex.batched_fc1_w.dtype != x.dtype or (layer.model_dim % 8) or (ex.hidden_size % 8) or (ex.output_dim % 8):
            return None
        if ex.output_dim != layer.model_dim and os.environ.get('TUTEL_B200_FUSED_OUTPUT_DIM', '0') != '1':
            return None     # `output_dim` experts: supported by the buffers/kernels, not yet covered by a multi-GPU test
    elif isinstance(ex, LlamaFFNNetwork):
        if ex.fp8 or G.classify_activation(ex.activation_fn) not in G.ACT_CODES or ex.W_fc1.dtype != x.dtype:
            return None
        if any(int(v) % 8 for v in ex.full_shapes['W_fc1'][1:]) or int(ex.full_shapes['W_fc3'][2]) != layer.model_dim:
            return None
    else:
        return None
    eng = layer.__dict__.get('_tb_fused_state', False)
    if eng is False:
        t = p2p.transport_for(layer.group)
        eng = FusedEngine(layer, t, x.dtype) if t is not None else None
        layer.__dict__['_tb_fused_state'] = eng        # kept on the layer itself (id() values get recycled)
    if eng is None or eng.dtype != x.dtype:
        return None
    plan = DispatchPlan.from_critical(crit)
    need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in layer.parameters()))
    try:
        bufs = eng.acquire(plan.C, hold=need_grad)
    except RuntimeError:
        bufs = None
    if bufs is None:
        if not eng.warned:
            import logging
            logging.warning('tutel_b200: symmetric heap too small for the fused MoE buffers of this layer (C=%d); using the '
                            'generic all-to-all path. Raise TUTEL_B200_HEAP_MB to enable the fused engine.', plan.C)
            eng.warned = True
        return None
    return _Runner(eng, d, plan, bufs)


class _Runner:
    def __init__(self, eng: FusedEngine, d: int, plan: DispatchPlan, bufs: _BufferSet):
        self.eng, self.d, self.plan, self.bufs = eng, d, plan, bufs

    def run(self, layer, x: torch.Tensor, crit) -> torch.Tensor:
        eng, plan, bufs = self.eng, self.plan, self.bufs
        ex = layer.experts
        gates = crit.gates_ks if hasattr(crit, 'gates_ks') else torch.stack([g.view(-1) for g in crit[3]])
        if hasattr(ex, 'full_shapes'):
            w1, w2, w3 = (ex._full(n, layer.group) for n in ('W_fc1', 'W_fc2', 'W_fc3'))
            return _FusedGLUMoE.apply(eng, bufs, plan, self.d, layer.is_postscore, G.classify_activation(ex.activation_fn),
                                      x, gates, w1, w2, w3)
        return _FusedMoE.apply(eng, bufs, plan, self.d, layer.is_postscore, x, gates, ex.batched_fc1_w,
                               ex.batched_fc1_bias, ex.batched_fc2_w, ex.batched_fc2_bias)


def _push(eng: FusedEngine, bufs: _BufferSet, plan: DispatchPlan, src: torch.Tensor, gates_f32, data_off: int,
          flag_off: int, width: int, chunk: int) -> torch.cuda.Event:
    """Launch the scatter-and-send kernel on the side stream; returns an event recorded after it."""
    C_ext = backend.require_ext()
    dst_tab, sig_tab = eng.push_tables(bufs, data_off, flag_off, width)
    cur = torch.cuda.current_stream()
    eng.side.wait_stream(cur)
    chunks_per_expert = -(-plan.C // chunk)
    with torch.cuda.stream(eng.side):
        backend.count_launch()
        C_ext.encode_rows(src, gates_f32, plan.slot_src, src, plan.k, plan.E, plan.C, dst_tab, sig_tab, chunk,
                          eng.rank * eng.El * chunks_per_expert, bufs.epoch, eng.chunk_counters.data_ptr())
        ev = torch.cuda.Event()
        ev.record(eng.side)
    src.record_stream(eng.side)
    plan.slot_src.record_stream(eng.side)
    if gates_f32 is not None:
        gates_f32.record_stream(eng.side)
    return ev


class _FusedMoE(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, eng: FusedEngine, bufs: _BufferSet, plan: DispatchPlan, d: int, is_postscore: bool,
                x, gates, w1, b1, w2, b2):
        t, W, El, E, rank = eng.t, eng.W, eng.El, eng.E, eng.rank
        C, M, H, Mo = plan.C, eng.M, eng.H, eng.Mo
        chunk = eng.chunk_rows(C, d)
        bufs.epoch += 1
        gates_f32 = gates.detach().to(torch.float32).contiguous()
        base = t.base_ptr(rank)

        # (1) dispatch: tokens -> expert GPUs' X_recv (side stream, overlaps GEMM1)
        ev = _push(eng, bufs, plan, x, None if is_postscore else gates_f32, bufs.x_recv, bufs.f_disp, M, chunk)

        # (2) GEMM1 + bias + ReLU on rows as they arrive
        x_recv = t.view(bufs.x_recv, [El * W, C, M], eng.dtype)
        cg, bn1, _ = eng.tile_counts(C, H)
        act = G.raw_gemm(x_recv, w1, epilogue=G.EPI_BIAS_RELU, bias=b1, b_group_div=W, cta_group=cg, block_n=bn1,
                         wait_flags=base + bufs.f_disp, wait_rows_per_flag=chunk, wait_flags_per_group=_FLAGS_PER_SEG,
                         wait_target=bufs.epoch, group_rot=rank, group_mod=-W)

        # (3) GEMM2 + bias, epilogue writes into the source GPUs' Y_comb and signals per expert
        cg2, bn2, tiles2 = eng.tile_counts(C, Mo)
        d_tab, s_tab = eng.combine_tables(bufs, bufs.y_comb, bufs.f_comb, Mo)
        y_comb = t.view(bufs.y_comb, [E, C, Mo], eng.dtype)
        G.raw_gemm(act, w2, b_mn=True, epilogue=G.EPI_BIAS if b2 is not None else G.EPI_NONE, bias=b2, b_group_div=W,
                   out=y_comb, cta_group=cg2, block_n=bn2, d_ptr_table=d_tab, signal_ptr_table=s_tab, group_rot=rank,
                   group_mod=-W)
        bufs.comb_total[0] += tiles2

        # (4) combine: weighted sum of each token's k rows once their experts have delivered
        backend.count_launch()
        out = backend.require_ext().decode_rows(y_comb.view(E * C, Mo), gates_f32 if is_postscore else None, plan.idx_ks,
                                                plan.loc_ks, E, C, base + bufs.f_comb, bufs.comb_total[0])
        torch.cuda.current_stream().wait_event(ev)

        ctx.eng, ctx.bufs, ctx.plan, ctx.d, ctx.is_postscore = eng, bufs, plan, d, is_postscore
        ctx.has_b1, ctx.has_b2 = b1 is not None, b2 is not None
        ctx.save_for_backward(x, gates, w1, w2, act)
        if not bufs.busy:
            pass  # inference: the set was never reserved
        return out

    @staticmethod
    def backward(ctx: Any, dout: torch.Tensor):
        eng, bufs, plan, d, is_postscore = ctx.eng, ctx.bufs, ctx.plan, ctx.d, ctx.is_postscore
        x, gates, w1, w2, act = ctx.saved_tensors
        t, W, El, E, rank = eng.t, eng.W, eng.El, eng.E, eng.rank
        C, M, H, Mo = plan.C, eng.M, eng.H, eng.Mo
        C_ext = backend.require_ext()
        chunk = eng.chunk_rows(C, d)
        base = t.base_ptr(rank)
        dout = dout.contiguous()
        gates_f32 = gates.detach().to(torch.float32).contiguous()
        x_recv = t.view(bufs.x_recv, [El * W, C, M], eng.dtype)
        y_comb = t.view(bufs.y_comb, [E * C, Mo], eng.dtype)

        # (a) gate gradients of the combine:  <dout[s], y[slot_j(s)]>
        dgates = None
        if is_postscore and ctx.needs_input_grad[6]:
            backend.count_launch()
            dgates = C_ext.gate_grad(dout, y_comb, plan.idx_ks, plan.loc_ks, E, C).to(gates.dtype)

        # (b) dispatch the output gradients to the expert GPUs (decode.bwd == encode of dout)
        bufs.epoch += 1
        ev = _push(eng, bufs, plan, dout, gates_f32 if is_postscore else None, bufs.dy_recv, bufs.b_disp, Mo, chunk)
        dy_recv = t.view(bufs.dy_recv, [El * W, C, Mo], eng.dtype)

        # (c) dh = (dy @ W2^T) * relu'(act)   as rows arrive
        cg, bnh, _ = eng.tile_counts(C, H)
        want_db1 = ctx.has_b1 and ctx.needs_input_grad[8]
        db1_acc = torch.zeros([El, H], dtype=torch.float32, device=dout.device) if want_db1 else None
        dh = G.raw_gemm(dy_recv, w2, epilogue=G.EPI_RELU_BWD, aux=act, b_group_div=W, cta_group=cg, block_n=bnh, colsum=db1_acc,
                        wait_flags=base + bufs.b_disp, wait_rows_per_flag=chunk, wait_flags_per_group=_FLAGS_PER_SEG,
                        wait_target=bufs.epoch, group_rot=rank, group_mod=-W)

        # (e) dX_e = dh @ W1, epilogue pushes into the source GPUs' dX_comb.  Collective decision: every rank
        #     must run it if any rank needs input gradients - the flag is part of the saved context (same program).
        need_dx = ctx.needs_input_grad[5] or (not is_postscore and ctx.needs_input_grad[6])
        dx_comb = t.view(bufs.dx_comb, [E, C, M], eng.dtype)
        if need_dx:
            cgx, bnx, tilesx = eng.tile_counts(C, M)
            d_tab, s_tab = eng.combine_tables(bufs, bufs.dx_comb, bufs.b_comb, M)
            G.raw_gemm(dh, w1, b_mn=True, b_group_div=W, out=dx_comb, cta_group=cgx, block_n=bnx, d_ptr_table=d_tab,
                       signal_ptr_table=s_tab, group_rot=rank, group_mod=-W)
            bufs.comb_total[1] += tilesx

        # (d, f) weight gradients over all W*C received rows of each local expert
        act_e, dh_e = act.view(El, W * C, H), dh.view(El, W * C, H)
        dw2 = G.raw_gemm(act_e, dy_recv.view(El, W * C, Mo), a_mn=True, b_mn=True) if ctx.needs_input_grad[9] else None
        dw1 = G.raw_gemm(dh_e, x_recv.view(El, W * C, M), a_mn=True, b_mn=True) if ctx.needs_input_grad[7] else None
        db1 = db1_acc.to(dh.dtype) if want_db1 else None
        db2 = G.column_sums(dy_recv.view(El, W * C, Mo)) if ctx.has_b2 and ctx.needs_input_grad[10] else None

        # (g) combine the input gradients (encode.bwd == decode of the gradient buffer)
        dx = None
        if need_dx:
            backend.count_launch()
            dx = C_ext.decode_rows(dx_comb.view(E * C, M), None if is_postscore else gates_f32, plan.idx_ks, plan.loc_ks,
                                   E, C, base + bufs.b_comb, bufs.comb_total[1])
        if not is_postscore and ctx.needs_input_grad[6]:
            backend.count_launch()
            dgates = C_ext.gate_grad(x, dx_comb.view(E * C, M), plan.idx_ks, plan.loc_ks, E, C).to(gates.dtype)
        torch.cuda.current_stream().wait_event(ev)
        bufs.busy = False
        return None, None, None, None, None, dx, dgates, dw1, db1, dw2, db2


class _FusedGLUMoE(torch.autograd.Function):
    """The same engine for gated (SwiGLU / "LLaMA") experts: dispatch feeds the dual-B GLU GEMM tile by tile, the down
    projection writes into the source GPUs' combine buffers; backward mirrors it with the fused GLU-gradient epilogue.
    (reference: tutel/experts/llama_ffn.py:38-41 between the two all-to-alls of tutel/impls/moe_layer.py:349-351)"""

    @staticmethod
    def forward(ctx: Any, eng: FusedEngine, bufs: _BufferSet, plan: DispatchPlan, d: int, is_postscore: bool, act: str,
                x, gates, w1, w2, w3):
        t, W, El, E, rank = eng.t, eng.W, eng.El, eng.E, eng.rank
        C, M, H, Mo = plan.C, eng.M, eng.H, eng.Mo
        chunk = eng.chunk_rows(C, d)
        bufs.epoch += 1
        gates_f32 = gates.detach().to(torch.float32).contiguous()
        base = t.base_ptr(rank)
        need_grad = any(ctx.needs_input_grad[6:])
        ev = _push(eng, bufs, plan, x, None if is_postscore else gates_f32, bufs.x_recv, bufs.f_disp, M, chunk)
        x_recv = t.view(bufs.x_recv, [El * W, C, M], eng.dtype)
        cg, _, _ = eng.tile_counts(C, H)
        h, g, u = G.glu_gemm(x_recv, w1, w2, b_mn=True, act=act, save_pre=need_grad, b_group_div=W, cta_group=cg,
                             wait_flags=base + bufs.f_disp, wait_rows_per_flag=chunk, wait_flags_per_group=_FLAGS_PER_SEG,
                             wait_target=bufs.epoch, group_rot=rank, group_mod=-W)
        cg2, bn2, tiles2 = eng.tile_counts(C, Mo)
        d_tab, s_tab = eng.combine_tables(bufs, bufs.y_comb, bufs.f_comb, Mo)
        y_comb = t.view(bufs.y_comb, [E, C, Mo], eng.dtype)
        G.raw_gemm(h, w3, b_mn=True, b_group_div=W, out=y_comb, cta_group=cg2, block_n=bn2, d_ptr_table=d_tab,
                   signal_ptr_table=s_tab, group_rot=rank, group_mod=-W)
        bufs.comb_total[0] += tiles2
        backend.count_launch()
        out = backend.require_ext().decode_rows(y_comb.view(E * C, Mo), gates_f32 if is_postscore else None, plan.idx_ks,
                                                plan.loc_ks, E, C, base + bufs.f_comb, bufs.comb_total[0])
        torch.cuda.current_stream().wait_event(ev)
        ctx.eng, ctx.bufs, ctx.plan, ctx.d, ctx.is_postscore, ctx.act = eng, bufs, plan, d, is_postscore, act
        if need_grad:
            ctx.save_for_backward(x, gates, w1, w2, w3, g, u, h)
        return out

    @staticmethod
    def backward(ctx: Any, dout: torch.Tensor):
        eng, bufs, plan, d, is_postscore, act = ctx.eng, ctx.bufs, ctx.plan, ctx.d, ctx.is_postscore, ctx.act
        x, gates, w1, w2, w3, g, u, h = ctx.saved_tensors
        t, W, El, E, rank = eng.t, eng.W, eng.El, eng.E, eng.rank
        C, M, H, Mo = plan.C, eng.M, eng.H, eng.Mo
        C_ext = backend.require_ext()
        chunk = eng.chunk_rows(C, d)
        base = t.base_ptr(rank)
        dout = dout.contiguous()
        gates_f32 = gates.detach().to(torch.float32).contiguous()
        x_recv = t.view(bufs.x_recv, [El * W, C, M], eng.dtype)
        y_comb = t.view(bufs.y_comb, [E * C, Mo], eng.dtype)
        dgates = None
        if is_postscore and ctx.needs_input_grad[7]:
            backend.count_launch()
            dgates = C_ext.gate_grad(dout, y_comb, plan.idx_ks, plan.loc_ks, E, C).to(gates.dtype)
        bufs.epoch += 1
        ev = _push(eng, bufs, plan, dout, gates_f32 if is_postscore else None, bufs.dy_recv, bufs.b_disp, Mo, chunk)
        dy_recv = t.view(bufs.dy_recv, [El * W, C, Mo], eng.dtype)
        cg, _, _ = eng.tile_counts(C, H)
        # dh = dy @ W3^T stays in TMEM; the epilogue emits dg and du as the gradient rows arrive
        dg, du = G.glu_gemm_bwd(dy_recv, w3, g, u, b_mn=False, act=act, b_group_div=W, cta_group=cg,
                                wait_flags=base + bufs.b_disp, wait_rows_per_flag=chunk,
                                wait_flags_per_group=_FLAGS_PER_SEG, wait_target=bufs.epoch, group_rot=rank, group_mod=-W)
        need_dx = ctx.needs_input_grad[6] or (not is_postscore and ctx.needs_input_grad[7])
        dx_comb = t.view(bufs.dx_comb, [E, C, M], eng.dtype)
        if need_dx:
            cgx, bnx, tilesx = eng.tile_counts(C, M)
            d_tab, s_tab = eng.combine_tables(bufs, bufs.dx_comb, bufs.b_comb, M)
            part = G.raw_gemm(dg, w1, b_group_div=W, cta_group=cgx, block_n=bnx)           # dg @ W1^T (local)
            G.raw_gemm(du, w2, epilogue=G.EPI_ADD, aux=part, b_group_div=W, out=dx_comb, cta_group=cgx, block_n=bnx,
                       d_ptr_table=d_tab, signal_ptr_table=s_tab, group_rot=rank, group_mod=-W)
            bufs.comb_total[1] += tilesx
        x_e, dy_e = x_recv.view(El, W * C, M), dy_recv.view(El, W * C, Mo)
        dw3 = G.raw_gemm(h.view(El, W * C, H), dy_e, a_mn=True, b_mn=True) if ctx.needs_input_grad[10] else None
        dw1 = G.raw_gemm(x_e, dg.view(El, W * C, H), a_mn=True, b_mn=True) if ctx.needs_input_grad[8] else None
        dw2 = G.raw_gemm(x_e, du.view(El, W * C, H), a_mn=True, b_mn=True) if ctx.needs_input_grad[9] else None
        dx = None
        if need_dx:
            backend.count_launch()
            dx = C_ext.decode_rows(dx_comb.view(E * C, M), None if is_postscore else gates_f32, plan.idx_ks, plan.loc_ks,
                                   E, C, base + bufs.b_comb, bufs.comb_total[1])
        if not is_postscore and ctx.needs_input_grad[7]:
            backend.count_launch()
            dgates = C_ext.gate_grad(x, dx_comb.view(E * C, M), plan.idx_ks, plan.loc_ks, E, C).to(gates.dtype)
        torch.cuda.current_stream().wait_event(ev)
        bufs.busy = False
        return None, None, None, None, None, None, dx, dgates, dw1, dw2, dw3
