"""Teacher inference = forward + aggregation + decode, batched and device resident.

Mirrors what the reference's drivers do one image at a time:

* ``aggregate_intree``  -- validate_hhrnet.py:92-101 (bilinear, align_corners=True, to the
  original image size; heat-maps from the 1/2-resolution head, tags from the 1/4 head);
* ``aggregate_flip_multiscale`` -- the flip-test / multi-scale protocol of
  legacy/valid_ae_avg.py:159-205 (upstream get_multi_stage_outputs + aggregate_results,
  configuration legacy/distillation.py:85-92);
* ``TeacherPipeline`` -- model -> aggregation -> ``HeatmapParser`` for a whole batch with one
  host<->device round trip; ``shard``/``gather`` give the one-process-per-GPU data-parallel
  form (images are independent; the only collective is the final gather of the padded
  keypoint tensors over NCCL).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import os

import torch

from . import _lib as L

# COCO left/right swap for the joint order of teacher_inference.py:38-40
FLIP_INDEX = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def bilinear_resize(src, out_hw, align_corners, dst=None, dst_inner=1, dst_off=0):
    """F.interpolate(src, out_hw, mode='bilinear', align_corners=...) on a CUDA NCHW tensor."""
    lib = L.load()
    src = _f32c(src)
    n, c, hi, wi = src.shape
    ho, wo = int(out_hw[0]), int(out_hw[1])
    if dst is None:
        dst = torch.empty((n, c, ho, wo), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        L.check(lib.brtpe_bilinear_resize(L.ptr(src), hi * wi, n * c, hi, wi, L.ptr(dst), ho, wo,
                                          dst_inner, dst_off, int(bool(align_corners)),
                                          L.stream_ptr(src.device)), "brtpe_bilinear_resize")
    return dst


def aggregate_intree(y0, y1, out_hw, num_joints=17):
    """validate_hhrnet.py:94-101 -> det (N,J,h,w), tag (N,A,h,w,1)."""
    det = bilinear_resize(y1, out_hw, True)
    tags = _f32c(y0)[:, num_joints:].contiguous()
    tag = bilinear_resize(tags, out_hw, True)
    return det, tag.unsqueeze(-1)


def aggregate_scale(y0, y1, y0f, y1f, base_hw, num_joints=17, det=None, tag=None,
                    accumulate=False, final_div=0.0, want_tags=True, flip_index=FLIP_INDEX):
    """One scale of the flip / multi-scale aggregation (see brtpe_aggregate_scale)."""
    lib = L.load()
    y0, y1 = _f32c(y0), _f32c(y1)
    flip = y0f is not None
    if flip:
        y0f, y1f = _f32c(y0f), _f32c(y1f)
    n, c0, h4, w4 = y0.shape
    j = num_joints
    a = c0 - j
    h2, w2 = y1.shape[2], y1.shape[3]
    hb, wb = int(base_hw[0]), int(base_hw[1])
    dev = y0.device
    if det is None:
        det = torch.empty((n, j, hb, wb), dtype=torch.float32, device=dev)
    t = 2 if flip else 1
    if want_tags and tag is None:
        tag = torch.empty((n, a, hb, wb, t), dtype=torch.float32, device=dev)
    fi = (C.c_int32 * j)(*[int(v) for v in flip_index[:j]])
    with torch.cuda.device(dev):
        L.check(lib.brtpe_aggregate_scale(
            L.ptr(y0), L.ptr(y1), L.ptr(y0f) if flip else None, L.ptr(y1f) if flip else None,
            n, j, a, h4, w4, h2, w2, hb, wb, fi, int(bool(accumulate)), float(final_div),
            L.ptr(det), L.ptr(tag) if want_tags else None, L.stream_ptr(dev)),
            "brtpe_aggregate_scale")
    return det, (tag if want_tags else None)


def aggregate_flip_multiscale(per_scale_outputs, base_size, num_joints=17):
    """``per_scale_outputs``: [(scale, [y0, y1], [y0f, y1f] or None), ...] in visiting order
    (descending scales, legacy/valid_ae_avg.py:166); ``base_size`` = (W, H).
    -> det (N,J,H,W), tag (N,A,H,W,T).  Tags come from the scale-1.0 pass only."""
    hb, wb = int(base_size[1]), int(base_size[0])
    ns = len(per_scale_outputs)
    det, tag = None, None
    for k, (scale, outs, outs_f) in enumerate(per_scale_outputs):
        want_tags = (scale == 1) or ns == 1
        last = k == ns - 1
        det, tg = aggregate_scale(outs[0], outs[1], outs_f[0] if outs_f else None,
                                  outs_f[1] if outs_f else None, (hb, wb), num_joints, det=det,
                                  accumulate=k > 0, final_div=float(ns) if last else 0.0,
                                  want_tags=want_tags)
        if want_tags:
            tag = tg
    return det, tag


class TeacherPipeline:
    """forward (+ flip test) -> aggregation -> decode for a batch of equally sized images.

    ``model``: this package's network (optionally wrapped by ``network_to_half``);
    ``parser``: this package's ``HeatmapParser``."""

    def __init__(self, model, parser, flip_test=True, project_hw=None, mode="upstream"):
        if mode not in ("upstream", "intree"):
            raise ValueError("mode must be 'upstream' (flip/project) or 'intree' (validate_hhrnet.py)")
        self.model = model
        self.parser = parser
        self.flip_test = bool(flip_test)
        self.project_hw = project_hw
        self.mode = mode
        self.num_joints = parser.params.num_joints

    def _forward_flip(self, x):
        """network outputs of cat(x, flip(x)): through the fused flip-pair plan when the model is
        this package's network (optionally inside ``network_to_half``'s Sequential), else by
        materialising the batch."""
        net, via_half = self.model, False
        if isinstance(net, torch.nn.Sequential) and len(net) == 3 and hasattr(net[1], "forward_flip_pair"):
            net, via_half = net[1], True                  # tofp16 -> net -> tofp32
        if hasattr(net, "forward_flip_pair") and x.dtype in (torch.float32, torch.float16) and \
                net.supports_flip_pair(x) and not os.environ.get("BRTPE_NO_FLIP_PAIR"):
            # borrowed plan outputs: aggregation consumes them in stream order before the next forward
            y0, y1 = net.forward_flip_pair(x, via_half=via_half and x.dtype == torch.float32,
                                           borrow=True)
            return y0.float(), y1.float()
        return self.model(torch.cat((x, torch.flip(x, [3])), 0))

    @torch.no_grad()
    def forward_aggregate(self, x, after_forward=None):
        """x (N,3,H,W) CUDA -> det (N,J,Hb,Wb), tag (N,A,Hb,Wb,T).  ``after_forward`` (optional
        callable) runs on the host right after the network launches have been enqueued."""
        n, _, h, w = x.shape
        hb, wb = self.project_hw if self.project_hw is not None else (h, w)
        if self.mode == "intree":
            y0, y1 = self.model(x)
            if after_forward is not None:
                after_forward()
            return aggregate_intree(y0, y1, (hb, wb), self.num_joints)
        if self.flip_test:
            y0, y1 = self._forward_flip(x)
            if after_forward is not None:
                after_forward()
            det, tag = aggregate_scale(y0[:n], y1[:n], y0[n:], y1[n:], (hb, wb), self.num_joints)
        else:
            y0, y1 = self.model(x)
            if after_forward is not None:
                after_forward()
            det, tag = aggregate_scale(y0, y1, None, None, (hb, wb), self.num_joints)
        return det, tag

    @torch.no_grad()
    def forward_aggregate_multiscale(self, inputs_by_scale, base_hw):
        """Multi-scale test protocol of legacy/valid_ae_avg.py:159-205: ``inputs_by_scale`` is a
        list of ``(scale, x_scale)`` with x_scale (N,3,Hs,Ws) CUDA -- the image batch already
        resized for that scale (the reference does it with resize_align_multi_scale) --
        visited in the order given (the reference visits descending scales).  Every scale runs
        the network on the image and, with ``flip_test``, on its mirror; heat-maps are
        projected to ``base_hw`` = (H, W), summed over scales and divided by their number,
        tags come from the scale-1.0 pass.  -> det (N,J,H,W), tag (N,A,H,W,T)."""
        hb, wb = int(base_hw[0]), int(base_hw[1])
        ns = len(inputs_by_scale)
        det, tag = None, None
        for k, (scale, x) in enumerate(inputs_by_scale):
            n = x.shape[0]
            if self.flip_test:
                y0, y1 = self._forward_flip(x)
                outs, outs_f = (y0[:n], y1[:n]), (y0[n:], y1[n:])
            else:
                outs, outs_f = tuple(self.model(x)), (None, None)
            want_tags = (scale == 1) or ns == 1
            det, tg = aggregate_scale(outs[0], outs[1], outs_f[0], outs_f[1], (hb, wb),
                                      self.num_joints, det=det, accumulate=k > 0,
                                      final_div=float(ns) if k == ns - 1 else 0.0,
                                      want_tags=want_tags)
            if want_tags:
                tag = tg
        return det, tag

    @torch.no_grad()
    def run_device(self, x, adjust=True, refine=True):
        """-> ans (N,Pmax,J,3+T), count (N), scores (N,Pmax): CUDA tensors, Pmax = J*K (the
        reference's bound), NO host synchronisation."""
        det, tag = self.forward_aggregate(x)
        return self.parser.decode_device(det, tag, adjust, refine, full_capacity=True)

    @torch.no_grad()
    def run_stream(self, host_batches, adjust=True, refine=True, early_images=None):
        """Pipelined host-facing loop over an iterable of (pinned) host batches.  The copy of
        batch i+1 into the other of two persistent device buffers runs on a side stream while
        batch i is aggregated and decoded -- it is released when the network of batch i has
        finished, because a copy that runs concurrently with the network's CUDA graph slows the
        graph down by more than the copy takes (measured: profiles/r01e_halo_pair.md).
        ``early_images`` images of the next batch are released already when the network of batch i
        STARTS: aggregation + decode (2.3 ms) have become shorter than a 157 MB copy (3 ms), and a
        partial copy under the graph costs less than exposing the rest (measured r01n/r01o, e2e /
        device throughput with the fused flip batch: 16 of 32 early 0.985, 22 -> 0.992, 28 -> 0.999).
        Default: 7/8 of the batch (environment variable BRTPE_EARLY_IMAGES overrides).
        Yields the device results ``(ans, count, scores)`` of every batch in order (see
        ``run_device``); the caller copies what it needs back."""
        L.load()
        if early_images is None and "BRTPE_EARLY_IMAGES" in os.environ:
            early_images = int(os.environ["BRTPE_EARLY_IMAGES"])
        dev = torch.device("cuda", torch.cuda.current_device())
        main = torch.cuda.current_stream(dev)
        # the copy stream and the two device input buffers belong to the pipeline, not to one call
        # of this generator: a buffer handed back to the caching allocator could be given to other
        # main-stream work that is still pending when the next call's first side-stream copy lands
        if getattr(self, "_copy_stream", None) is None or self._copy_stream.device != dev:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._in_bufs = [None, None]
        copy_stream = self._copy_stream
        bufs = self._in_bufs
        # the first copy has no event to wait for: order the whole copy stream after everything
        # already enqueued on the main stream (the last readers of the buffers included)
        copy_stream.wait_stream(main)
        state = {"k": 0}

        def stage_part(k, xh, lo, hi, after):
            if hi <= lo:
                return None
            if after is not None:
                copy_stream.wait_event(after)  # also orders the copy after the last reader of bufs[k]
            with torch.cuda.stream(copy_stream):
                bufs[k][lo:hi].copy_(xh[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return ev

        def claim(xh):
            k = state["k"]
            state["k"] = k ^ 1
            if bufs[k] is None or bufs[k].shape != xh.shape or bufs[k].dtype != xh.dtype:
                bufs[k] = torch.empty(xh.shape, dtype=xh.dtype, device=dev)
            return k

        it = iter(host_batches)
        try:
            x0 = next(it)
        except StopIteration:
            return
        k0 = claim(x0)
        nxt = [(k0, [stage_part(k0, x0, 0, x0.shape[0], None)])]
        while nxt[0] is not None:
            k, evs = nxt[0]
            for ev in evs:
                if ev is not None:
                    main.wait_event(ev)
            pending = {}
            try:
                xn = next(it)
            except StopIteration:
                xn = None
            if xn is not None:
                kn = claim(xn)
                ne = (xn.shape[0] * 7) // 8 if early_images is None else int(early_images)
                ne = min(max(ne, 0), xn.shape[0])
                start = torch.cuda.Event()
                start.record(main)             # after the previous step's last reader of bufs[kn]
                pending = {"k": kn, "x": xn, "ne": ne, "evs": [stage_part(kn, xn, 0, ne, start)]}

            def stage_next():
                if not pending:
                    nxt[0] = None
                    return
                fwd_done = torch.cuda.Event()
                fwd_done.record(main)
                xq = pending["x"]
                pending["evs"].append(stage_part(pending["k"], xq, pending["ne"], xq.shape[0], fwd_done))
                nxt[0] = (pending["k"], pending["evs"])

            det, tag = self.forward_aggregate(bufs[k], after_forward=stage_next)
            yield self.parser.decode_device(det, tag, adjust, refine, full_capacity=True)

    @torch.no_grad()
    def run(self, x, adjust=True, refine=True):
        """Host-facing call: ``x`` may live in (pinned) host memory.  -> list over images of
        (people (P,J,3+T) float32, scores [np.float32])."""
        L.load()
        dev = torch.device("cuda", torch.cuda.current_device())
        xd = x.to(dev, non_blocking=True)
        ans, count, scores = self.run_device(xd, adjust, refine)
        return unpack_results(ans.cpu(), count.cpu(), scores.cpu())


def unpack_results(ans, count, scores):
    ans, count, scores = ans.numpy(), count.numpy(), scores.numpy()
    out = []
    for i in range(ans.shape[0]):
        c = int(count[i])
        people = np.array(ans[i, :c]) if c else np.array([], dtype=np.float32)
        out.append((people, [np.float32(s) for s in scores[i, :c]]))
    return out


# ---------------------------------------------------------------------------------------
# data-parallel sharding: one process per GPU, images are independent
# ---------------------------------------------------------------------------------------
def shard_range(total, rank, world):
    """Contiguous split of ``total`` images over ``world`` ranks -> (start, stop)."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pad_results(ans, count, scores, pcap):
    """Fixed-shape payload for the gather: people beyond ``pcap`` are flagged, not dropped
    silently (``overflow`` = count > pcap)."""
    n, pmax = ans.shape[0], ans.shape[1]
    if pmax >= pcap:
        a, s = ans[:, :pcap], scores[:, :pcap]
    else:
        a = torch.zeros((n, pcap) + tuple(ans.shape[2:]), dtype=ans.dtype, device=ans.device)
        a[:, :pmax] = ans
        s = torch.zeros((n, pcap), dtype=scores.dtype, device=scores.device)
        s[:, :pmax] = scores
    return a.contiguous(), count.contiguous(), s.contiguous()


def pack_results(ans, count, scores):
    """(ans (N,P,J,3+T) f32, count (N) i32, scores (N,P) f32) -> ONE float32 payload (N, 1 + P +
    P*J*(3+T)) per image: [count (the int32 bit pattern) | scores | people], so that the final
    gather is a single collective."""
    n = ans.shape[0]
    return torch.cat((count.contiguous().view(torch.float32).view(n, 1), scores.reshape(n, -1),
                      ans.reshape(n, -1)), dim=1).contiguous()


def unpack_payload(payload, pcap, num_joints, width):
    """inverse of ``pack_results`` -> (ans, count, scores) views of the payload."""
    n = payload.shape[0]
    count = payload[:, 0].contiguous().view(torch.int32)
    scores = payload[:, 1:1 + pcap]
    ans = payload[:, 1 + pcap:].view(n, pcap, num_joints, width)
    return ans, count, scores


class ResultGatherer:
    """The only collective of the data-parallel path: ONE all-gather of the packed
    per-image results per batch, issued on a side stream so that it overlaps the next batch's
    network (NCCL on GPUs; gloo in the CPU tests, where streams do not exist).  The gathered tensor
    stays where it is -- nobody copies it to the host implicitly; every rank keeps (and may copy
    out) its own shard."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.stream = None
        self._out = {}

    def _buffer(self, payload, slot):
        key = (slot, tuple(payload.shape), payload.dtype, payload.device)
        buf = self._out.get(key)
        if buf is None:
            buf = torch.empty((self.world * payload.shape[0],) + tuple(payload.shape[1:]),
                              dtype=payload.dtype, device=payload.device)
            self._out[key] = buf
        return buf

    def gather(self, payload, slot=0):
        """-> (gathered (world*N, ...) in rank order, event or None).  On CUDA the collective runs
        on the gatherer's side stream after everything enqueued on the current stream so far; the
        caller waits for the returned event (``torch.cuda.current_stream().wait_event``) before it
        reads ``gathered`` or re-uses ``payload`` / the same ``slot``."""
        out = self._buffer(payload, slot)
        if not payload.is_cuda:
            self.dist.all_gather_into_tensor(out, payload.contiguous(), group=self.group)
            return out, None
        dev = payload.device
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=dev)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        payload.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.dist.all_gather_into_tensor(out, payload, group=self.group)
            done = torch.cuda.Event()
            done.record(self.stream)
        return out, done


def gather_results(ans, count, scores, group=None):
    """Blocking form: one packed all-gather of equally shaped per-rank results.
    -> concatenated (ans, count, scores) in rank order on every rank."""
    g = ResultGatherer(group)
    payload = pack_results(ans, count, scores)
    out, done = g.gather(payload)
    if done is not None:
        torch.cuda.current_stream(payload.device).wait_event(done)
    return unpack_payload(out, ans.shape[1], ans.shape[2], ans.shape[3])


def pin_to_gpu_numa(device_index):
    """Bind this process to the CPU cores next to its GPU (NVML's ideal CPU affinity = the GPU's
    NUMA node): eight ranks that each stream 157 MB per step from pinned host memory should not
    cross the socket interconnect.  Best effort: returns the number of cores bound, or 0."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cores = [c for c in cores if c in allowed]
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:
        pass
    return 0
