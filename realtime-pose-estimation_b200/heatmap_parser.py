"""``HeatmapParser`` -- drop-in for rtpe/third_party/group.py:125-287, on the GPU.

Same constructor, same methods (``nms``, ``top_k``, ``match``, ``adjust``, ``refine``,
``parse``), same argument meaning, same return structures and dtypes as the reference;
the work is done by the CUDA kernels of libbrtpe.so (csrc/decode_*.cu) through the C
ABI in include/brtpe.h.  ``parse_batch`` / ``decode_device`` are the batched entry
points ("reference parse applied to every image") that keep
everything on the device until one final copy.

Tie rule of top-k (undefined in the reference, group.py:153): value descending, flat
index ascending.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


class Params(object):
    """Mirror of group.py:100-122."""

    def __init__(self, num_joints, max_num_people, detection_threshold,
                 tag_threshold, use_detection_val, ignore_too_much):
        self.num_joints = num_joints
        self.max_num_people = max_num_people
        self.detection_threshold = detection_threshold
        self.tag_threshold = tag_threshold
        self.use_detection_val = use_detection_val
        self.ignore_too_much = ignore_too_much
        self.joint_order = list(range(num_joints))


def _as_cuda_f32(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.detach().to(device=device, dtype=torch.float32).contiguous()


class HeatmapParser(object):
    def __init__(self, num_joints, max_num_people, detection_threshold,
                 tag_threshold, use_detection_val, ignore_too_much,
                 tag_per_joint=True, nms_ksize=5, nms_padding=2,
                 munkres_start_rule="previous", person_capacity=128):
        self.params = Params(num_joints, max_num_people, detection_threshold,
                             tag_threshold, use_detection_val, ignore_too_much)
        self.tag_per_joint = tag_per_joint
        self.nms_ksize = nms_ksize
        self.nms_padding = nms_padding
        if munkres_start_rule not in ("previous", "origin"):
            raise ValueError("munkres_start_rule must be 'previous' (munkres 1.1.x) or 'origin'")
        self.munkres_start_rule = munkres_start_rule
        # first-try capacity of the per-image person list; the reference's list is
        # unbounded (<= J*K), so an overflow transparently re-runs with J*K.
        self.person_capacity = int(person_capacity)
        # after an overflow the following decodes start at the J*K bound right away (a stream of
        # crowded images -- e.g. an untrained student, ~170 "persons" per image -- would otherwise
        # run the grouping twice per batch); results do not depend on the capacity
        self._capacity_hint = 0
        self._ws = {}

    # ------------------------------------------------------------------ helpers
    def _device(self, t=None):
        L.load()
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
        return torch.device("cuda", torch.cuda.current_device())

    def _cparams(self):
        p = self.params
        return L.DecodeParams(int(p.num_joints), int(p.max_num_people),
                              float(p.detection_threshold), float(p.tag_threshold),
                              int(bool(p.use_detection_val)), int(bool(p.ignore_too_much)),
                              int(bool(self.tag_per_joint)), int(self.nms_ksize),
                              int(self.nms_padding),
                              0 if self.munkres_start_rule == "previous" else 1)

    def _workspace(self, name, nbytes, device):
        key = (name, str(device))
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._ws[key] = buf
        return buf

    def _pmax_full(self):
        return int(self.params.num_joints) * int(self.params.max_num_people)

    # ------------------------------------------------------------------ nms
    def nms(self, det):
        """group.py:134-138.  det (N,J,H,W) -> same shape, CUDA float32."""
        lib = L.load()
        dev = self._device(det)
        d = _as_cuda_f32(det, dev)
        n, j, h, w = d.shape
        out = torch.empty_like(d)
        with torch.cuda.device(dev):
            L.check(lib.brtpe_nms(L.ptr(d), L.ptr(out), n * j, h, w, self.nms_ksize,
                                  self.nms_padding, L.stream_ptr(dev)), "brtpe_nms")
        return out

    # ------------------------------------------------------------------ top_k
    def top_k_device(self, det, tag):
        """Fused NMS + top-K + tag gather on the device.
        -> val_k (N,J,K) f32, ind_k (N,J,K) i32, loc_k (N,J,K,2) i64, tag_k (N,J,K,T) f32."""
        lib = L.load()
        dev = det.device
        n, j, h, w = det.shape
        k = int(self.params.max_num_people)
        jt = tag.shape[1]
        t = tag.numel() // (tag.shape[0] * jt * h * w)
        if tag.shape[0] != n or tag.numel() != n * jt * h * w * t or jt not in (1, j):
            raise ValueError("tag shape %s does not match det shape %s"
                             % (tuple(tag.shape), tuple(det.shape)))
        if self.tag_per_joint and jt != j:
            raise ValueError("tag_per_joint=True needs %d tag planes, got %d" % (j, jt))
        if not self.tag_per_joint and jt != 1:
            # reference: tag.expand(-1, J, -1, -1) only works for a singleton dim
            raise ValueError("tag_per_joint=False needs exactly one tag plane, got %d" % jt)
        val_k = torch.empty((n, j, k), dtype=torch.float32, device=dev)
        ind_k = torch.empty((n, j, k), dtype=torch.int32, device=dev)
        loc_k = torch.empty((n, j, k, 2), dtype=torch.int64, device=dev)
        tag_k = torch.empty((n, j, k, t), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            wsb = lib.brtpe_topk_workspace_bytes(n, j, h, w, k)
            ws = self._workspace("topk", wsb, dev)
            L.check(lib.brtpe_nms_topk_gather(
                L.ptr(det), L.ptr(tag), n, j, jt, h, w, t, k, self.nms_ksize, self.nms_padding,
                L.ptr(val_k), L.ptr(ind_k), L.ptr(loc_k), L.ptr(tag_k), L.ptr(ws),
                ws.numel(), L.stream_ptr(dev)), "brtpe_nms_topk_gather")
        return val_k, ind_k, loc_k, tag_k

    def top_k(self, det, tag):
        """group.py:144-179 -> {'tag_k','loc_k','val_k'} numpy arrays."""
        dev = self._device(det)
        d = _as_cuda_f32(det, dev)
        t = _as_cuda_f32(tag, dev)
        val_k, _, loc_k, tag_k = self.top_k_device(d, t)
        return {"tag_k": tag_k.cpu().numpy(), "loc_k": loc_k.cpu().numpy(),
                "val_k": val_k.cpu().numpy()}

    # ------------------------------------------------------------------ match
    def match_device(self, val_k, ind_k, tag_k, width, pmax=None, defer_overflow=False):
        """Grouping on the device -> ans (N,Pmax,J,3+T) f32, count (N) i32, Pmax used.
        ``defer_overflow``: do not wait for the capacity flag; returns (ans, count, pmax, flag) with
        flag = the device int32 the caller must read later (None when pmax is already the bound)."""
        lib = L.load()
        dev = val_k.device
        n, j, k = val_k.shape
        t = tag_k.shape[3]
        if j != self.params.num_joints:
            raise ValueError("num_joints mismatch: parser %d, input %d" % (self.params.num_joints, j))
        if k != self.params.max_num_people:
            raise ValueError("K mismatch: parser max_num_people %d, input %d"
                             % (self.params.max_num_people, k))
        pmax_full = self._pmax_full()
        pmax = min(pmax_full, max(self.person_capacity, self._capacity_hint)) if pmax is None \
            else int(pmax)
        prm = self._cparams()
        with torch.cuda.device(dev):
            while True:
                ans = torch.empty((n, pmax, j, 3 + t), dtype=torch.float32, device=dev)
                count = torch.empty((n,), dtype=torch.int32, device=dev)
                overflow = torch.zeros((1,), dtype=torch.int32, device=dev)
                wsb = lib.brtpe_group_workspace_bytes(n, j, k, t, pmax)
                ws = self._workspace("group", wsb, dev)
                L.check(lib.brtpe_group_ae(L.ptr(val_k), L.ptr(ind_k), L.ptr(tag_k), n, int(width),
                                           t, C.byref(prm), L.ptr(ans), L.ptr(count),
                                           L.ptr(overflow), pmax, L.ptr(ws), ws.numel(),
                                           L.stream_ptr(dev)), "brtpe_group_ae")
                if defer_overflow:
                    return ans, count, pmax, (overflow if pmax < pmax_full else None)
                if pmax >= pmax_full or int(overflow.item()) == 0:
                    return ans, count, pmax
                pmax = self._capacity_hint = pmax_full

    def match(self, tag_k, loc_k, val_k):
        """group.py:140-142 -> list (one per image) of (P,J,3+T) float32 arrays."""
        dev = self._device()
        loc = np.asarray(loc_k)
        wfake = int(loc[..., 0].max()) + 1 if loc.size else 1
        ind = (loc[..., 1] * wfake + loc[..., 0]).astype(np.int32)
        v = _as_cuda_f32(np.asarray(val_k, np.float32), dev)
        tg = _as_cuda_f32(np.asarray(tag_k, np.float32), dev)
        i = torch.from_numpy(ind).to(dev).contiguous()
        ans, count, _ = self.match_device(v, i, tg, wfake)
        return self._unpack(ans, count)

    @staticmethod
    def _unpack(ans, count):
        ans = ans.cpu().numpy()
        count = count.cpu().numpy()
        out = []
        for n in range(ans.shape[0]):
            c = int(count[n])
            out.append(np.array(ans[n, :c]) if c > 0 else np.array([], dtype=np.float32))
        return out

    # ------------------------------------------------------------------ adjust
    def adjust_device(self, ans, count, det):
        lib = L.load()
        n, pmax, j, width = ans.shape
        with torch.cuda.device(det.device):
            L.check(lib.brtpe_adjust(L.ptr(ans), L.ptr(count), L.ptr(det), n, j, det.shape[2],
                                     det.shape[3], width - 3, pmax, L.stream_ptr(det.device)),
                    "brtpe_adjust")
        return ans

    def _pack(self, ans_list, dev):
        n = len(ans_list)
        counts = [int(a.shape[0]) if getattr(a, "size", 0) > 0 else 0 for a in ans_list]
        pmax = max(1, max(counts) if counts else 1)
        width = None
        for a in ans_list:
            if getattr(a, "size", 0) > 0:
                width = a.shape[2]
                j = a.shape[1]
        if width is None:
            return None, None, counts
        buf = np.zeros((n, pmax, j, width), np.float32)
        for i, a in enumerate(ans_list):
            if counts[i]:
                buf[i, :counts[i]] = a
        return (torch.from_numpy(buf).to(dev),
                torch.tensor(counts, dtype=torch.int32, device=dev), counts)

    def adjust(self, ans, det):
        """group.py:181-200; ``ans`` (list of per-image arrays) is updated in place."""
        dev = self._device(det)
        d = _as_cuda_f32(det, dev)
        buf, count, counts = self._pack(ans, dev)
        if buf is None:
            return ans
        self.adjust_device(buf, count, d)
        res = buf.cpu().numpy()
        for i, c in enumerate(counts):
            if c:
                ans[i][...] = res[i, :c]
        return ans

    # ------------------------------------------------------------------ refine
    def refine_device(self, det, tag, ans, count):
        lib = L.load()
        dev = det.device
        n, j, h, w = det.shape
        jt = tag.shape[1]
        t = tag.numel() // (n * jt * h * w)
        pmax = ans.shape[1]
        with torch.cuda.device(dev):
            wsb = lib.brtpe_refine_workspace_bytes(n, j, t, pmax)
            ws = self._workspace("refine", wsb, dev)
            L.check(lib.brtpe_refine(L.ptr(det), L.ptr(tag), L.ptr(ans), L.ptr(count), n, j, jt,
                                     h, w, t, pmax, L.ptr(ws), ws.numel(), L.stream_ptr(dev)),
                    "brtpe_refine")
        return ans

    def refine(self, det, tag, keypoints):
        """group.py:202-264 for ONE person: det (J,H,W), tag (J,H,W[,T]), keypoints (J,3+T)
        float32 (updated in place and returned)."""
        dev = self._device()
        det = np.asarray(det, np.float32)
        tag = np.asarray(tag, np.float32)
        if tag.ndim == 3:
            tag = tag[:, :, :, None]
        d = _as_cuda_f32(det[None], dev)
        tg = _as_cuda_f32(tag[None], dev)
        kp = torch.from_numpy(np.ascontiguousarray(keypoints, np.float32)[None, None]).to(dev)
        count = torch.ones((1,), dtype=torch.int32, device=dev)
        self.refine_device(d, tg, kp, count)
        keypoints[...] = kp.cpu().numpy()[0, 0]
        return keypoints

    # ------------------------------------------------------------------ parse
    def decode_device(self, det, tag, adjust=True, refine=True, full_capacity=False):
        """Whole decode on the device.
        -> ans (N,Pmax,J,3+T) f32, count (N) i32, scores (N,Pmax) f32 (all CUDA).
        ``full_capacity=False``: the person lists are sized ``person_capacity`` and the capacity
        flag is read once at the end (ONE host sync per decode; an overflow re-runs with J*K).
        ``full_capacity=True``: the lists are sized J*K = the reference's own bound (173 KB per image
        at T = 2), nothing can overflow and the call never synchronises with the host -- what the
        device-resident pipeline uses, so that decode(i) can be enqueued behind forward(i+1)."""
        lib = L.load()
        dev = det.device
        n, j, h, w = det.shape
        val_k, ind_k, _, tag_k = self.top_k_device(det, tag)
        if full_capacity:
            ans, count, pmax = self.match_device(val_k, ind_k, tag_k, w, pmax=self._pmax_full())
            flag = None
        else:
            # optimistic: everything is enqueued for the default person capacity and the capacity
            # flag is read once at the end (after the last launch, instead of a sync in the middle
            # that left the GPU waiting for the remaining launches)
            ans, count, pmax, flag = self.match_device(val_k, ind_k, tag_k, w, defer_overflow=True)
        while True:
            if adjust:
                self.adjust_device(ans, count, det)
            scores = torch.empty((n, pmax), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                L.check(lib.brtpe_scores(L.ptr(ans), L.ptr(count), L.ptr(scores), n, j,
                                         ans.shape[3] - 3, pmax, L.stream_ptr(dev)), "brtpe_scores")
            if refine:
                self.refine_device(det, tag, ans, count)
            if flag is None or int(flag.item()) == 0:
                return ans, count, scores
            self._capacity_hint = self._pmax_full()
            ans, count, pmax = self.match_device(val_k, ind_k, tag_k, w, pmax=self._pmax_full())
            flag = None

    def parse_batch(self, det, tag, adjust=True, refine=True):
        """Reference ``parse`` applied to every image of the batch.
        -> list of (people, scores): people (P,J,3+T) float32 array (shape (0,) if none),
        scores list of np.float32."""
        dev = self._device(det)
        d = _as_cuda_f32(det, dev)
        t = _as_cuda_f32(tag, dev)
        ans, count, scores = self.decode_device(d, t, adjust, refine)
        people = self._unpack(ans, count)
        sc = scores.cpu().numpy()
        return [(people[i], [np.float32(s) for s in sc[i, :people[i].shape[0]]]
                 if people[i].size else [])
                for i in range(len(people))]

    def parse(self, det, tag, adjust=True, refine=True):
        """group.py:266-287.  Like the reference, ``scores`` and (with ``refine=True``) the
        returned people are those of image 0 only."""
        dev = self._device(det)
        d = _as_cuda_f32(det, dev)
        t = _as_cuda_f32(tag, dev)
        if refine:
            d, t = d[:1], t[:1]
        ans, count, scores = self.decode_device(d.contiguous(), t.contiguous(), adjust, refine)
        people = self._unpack(ans, count)
        n0 = people[0].shape[0] if people[0].size else 0
        sc = [np.float32(s) for s in scores[0, :n0].cpu().numpy()]
        if refine:
            return [people[0]], sc
        return people, sc
