"""Data-parallel gradient exchange for the TBI_ResNest step (reference: tf.distribute.MirroredStrategy,
MainParallel.py:16,130 -- replicated variables, one all-reduce(SUM) of the gradients per step).

One process per GPU; parameters, Adam state and the flat gradient buffer are replicated; the batch is
sharded (weak scaling).  The only collective is the all-reduce of the flat fp32 gradient buffer, cut
into contiguous buckets in BACKWARD order and launched on a side stream as soon as the backward
program has finalised a bucket, so NCCL over NVLink/NVSwitch overlaps the remaining backward kernels.
Adam then applies grad_scale = 1/world (average).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def plan_buckets(marks: Sequence[Tuple[int, int]], total: int, bucket_elems: int) -> List[Tuple[int, int, int]]:
    """marks: (calls_issued, watermark) in backward order, watermark non-increasing; everything at flat
    offset >= watermark is final once `calls_issued` backward calls have been enqueued.
    -> [(calls_issued, lo, hi)]: after that many calls, all-reduce flat[lo:hi].  Buckets tile [0,total)
    exactly once, each >= bucket_elems except possibly the last."""
    out: List[Tuple[int, int, int]] = []
    hi = total
    for calls, wm in marks:
        wm = min(wm, hi)
        if hi - wm >= bucket_elems or (wm == 0 and hi > 0):
            out.append((calls, wm, hi))
            hi = wm
    if hi > 0:                                   # marks never reached 0 (should not happen): flush at the end
        out.append((marks[-1][0] if marks else 0, 0, hi))
    return out


def plan_buckets_at(marks: Sequence[Tuple[int, int]], total: int, cuts: Sequence[int]) -> List[Tuple[int, int, int]]:
    """Buckets with boundaries at the given flat offsets (descending = backward order): bucket i is all-reduced as soon as the
    backward program has finalised every gradient at offset >= cuts[i].  Same output format as plan_buckets."""
    out: List[Tuple[int, int, int]] = []
    hi = total
    want = sorted((c for c in set(cuts) if 0 < c < total), reverse=True)
    for calls, wm in marks:
        wm = min(wm, hi)
        while want and wm <= want[0]:
            lo = want.pop(0)
            if hi > lo:
                out.append((calls, lo, hi)); hi = lo
        if wm == 0 and hi > 0:
            out.append((calls, 0, hi)); hi = 0
    if hi > 0:
        out.append((marks[-1][0] if marks else 0, 0, hi))
    return out


class GradSync:
    """average=True (default): Adam sees the MEAN of the per-replica gradients, which is what the reference's data-parallel
    driver computes -- each replica's loss is divided by the global batch (compute_average_loss, VisionTransformer.py:225-227)
    and MirroredStrategy SUMS the replica gradients in apply_gradients (MainParallel.py:130).  average=False applies the raw
    SUM (what MirroredStrategy would do to TBI_ResNest.step's un-normalised [H,W] loss)."""

    def __init__(self, process_group=None, bucket_bytes: Optional[int] = None, average: bool = True):
        if bucket_bytes is None:                             # TBI_BUCKET_MB: sweep knob for bench runs
            # 54 MB = two buckets for the ~108 MB of fp32 gradients: every bucket boundary also joins the weight-gradient side
            # stream of the backward (engine.run_bwd), so fewer, larger buckets keep more of that overlap
            # (2 x B200: 9.93 ms/step with 25 MB buckets, 9.79 ms with 54 MB; 1 GPU: 9.61 ms)
            bucket_bytes = int(os.environ.get("TBI_BUCKET_MB", "54")) << 20
        if not dist.is_initialized():
            raise RuntimeError("GradSync needs torch.distributed to be initialised (backend nccl on GPUs)")
        self.pg = process_group
        self.world_size = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.average = average
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.comm_stream: Optional[torch.cuda.Stream] = None
        self._plan = None
        self._plan_key = None
        self._seg_graphs = {}

    def attach(self, engine):
        """replicated variables start identical on every replica (MirroredStrategy mirrors the creating replica's values): rank 0's
        parameters, moving statistics and Adam state are broadcast; every replica then draws its own dropout masks (independent
        noise per shard, as independent tf.nn.dropout calls per replica give)."""
        for t in (engine.params, engine.stats, engine.adam_m, engine.adam_v, engine.step_count):
            dist.broadcast(t, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
        engine.dropout_seed = 0x5EED + 1000003 * self.rank

    def allreduce(self, flat_slice: torch.Tensor):
        dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, group=self.pg)

    def gather(self, t: torch.Tensor) -> torch.Tensor:
        """eval-side counterpart of mirrored_strategy.gather(values, axis=0) (MainParallel.py:160,163): every replica's
        per-shard tensor (class scores [B/world, H, W, C], labels) concatenated along the batch axis, on every rank"""
        t = t.contiguous()
        parts = [torch.empty_like(t) for _ in range(self.world_size)]
        dist.all_gather(parts, t, group=self.pg)
        return torch.cat(parts, dim=0)

    def reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        """mirrored_strategy.reduce(SUM, per_replica_value, axis=None) for the loss / metric scalars (MainParallel.py:131-134,159)"""
        out = t.detach().clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.pg)
        return out

    def backward_and_sync(self, engine):
        """run engine.prog_bwd, interleaving bucket all-reduces on a side stream"""
        self._ensure_plan(engine)
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=engine.device)
        cur = torch.cuda.current_stream(engine.device)
        engine.grads.zero_()
        pos = 0
        for calls, lo, hi in self._plan:
            engine.run_bwd(pos, calls)
            pos = calls
            ev = torch.cuda.Event()
            ev.record(cur)
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                self.allreduce(engine.grads[lo:hi])
        engine.run_bwd(pos, len(engine.prog_bwd))
        cur.wait_stream(self.comm_stream)

    # ------------------------------------------------------------------ CUDA-graph replay with eager collectives
    def step_graphed(self, engine, pre_fn, post_fn, key, loss_fn=None, before_loss=None):
        """One training step as CUDA-graph SEGMENTS: [pre_fn = prepare/forward/loss] , one graph per backward slice
        between bucket boundaries, [post_fn = Adam].  The bucket all-reduces are launched eagerly on the side stream
        between segment replays (NCCL calls are never captured), so they overlap the following backward segments
        exactly as in the eager path.  With loss_fn the loss is its own segment and before_loss() runs (eagerly) in front of
        it: the caller's label copy joins there instead of in front of the forward pass."""
        self._ensure_plan(engine)
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=engine.device)
        cur = torch.cuda.current_stream(engine.device)
        graphs = self._seg_graphs.get(key)
        if graphs is None:
            torch.cuda.synchronize()

            def cap(fn):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                return g

            def pre():
                pre_fn()
                if loss_fn is None:
                    engine.grads.zero_()

            def loss_seg():
                loss_fn()
                engine.grads.zero_()

            segs, pos = [], 0
            for calls, lo, hi in self._plan:
                a, b = pos, calls
                segs.append((cap(lambda a=a, b=b: engine.run_bwd(a, b)) if b > a else None, lo, hi))
                pos = calls
            tail = cap(lambda: engine.run_bwd(pos, len(engine.prog_bwd))) if pos < len(engine.prog_bwd) else None
            graphs = dict(pre=cap(pre), loss=cap(loss_seg) if loss_fn is not None else None, segs=segs, tail=tail, post=cap(post_fn))
            self._seg_graphs[key] = graphs
            cur = torch.cuda.current_stream(engine.device)
        graphs["pre"].replay()
        if before_loss is not None:
            before_loss()
        if graphs["loss"] is not None:
            graphs["loss"].replay()
        for g, lo, hi in graphs["segs"]:
            if g is not None:
                g.replay()
            ev = torch.cuda.Event()
            ev.record(cur)
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                self.allreduce(engine.grads[lo:hi])
        if graphs["tail"] is not None:
            graphs["tail"].replay()
        cur.wait_stream(self.comm_stream)
        graphs["post"].replay()

    def _ensure_plan(self, engine):
        key = (id(engine), engine.gen)
        if self._plan_key != key:
            cuts = getattr(engine, "bucket_cuts", None)
            if cuts and os.environ.get("TBI_BUCKET_MB") is None:
                # boundaries chosen on the backward's TIMELINE (Engine.bucket_cuts, measured there): the decoder's transposed convs
                # and the two deepest encoder stages hold ~95 % of the gradient bytes and are final ~70 % of the way through the
                # backward; what is left for the end (the wide, slow, nearly parameter-free full-resolution layers) is a few
                # MB, so the exposed tail all-reduce is short
                self._plan = plan_buckets_at(engine.bwd_marks, engine.P.total, cuts)
            else:
                self._plan = plan_buckets(engine.bwd_marks, engine.P.total, self.bucket_elems)
            self._plan_key = key
