"""Benchmark of the teacher-inference hot path (BASELINE.json): HigherHRNet-W48 640x640
forward with flip test + heat-map aggregation + HeatmapParser decode.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference            # the CPU implementation of the same path

A step = one batch of 32 synthetic 640x640 images per GPU (configs[1] of BASELINE.json):
64 forwards (flip test), aggregation to 640x640 (T = 2 tag copies), parse(adjust, refine).
Prints ONE JSON line (rank 0).  `value` is timed with the inputs already in HBM; `e2e` goes
through the host-facing call with pinned host input and host results (H2D + D2H inside the
timed region).  Device timing = CUDA events on the launching stream, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
# banner to stdout at communicator creation), so the process's fd 1 is pointed at stderr and
# the JSON line goes to a private duplicate of the original stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


METRIC = "images/sec W48 640^2 fwd+decode"
PARSER_KW = dict(num_joints=17, max_num_people=30, detection_threshold=0.1, tag_threshold=1.0,
                 use_detection_val=True, ignore_too_much=False, tag_per_joint=True, nms_ksize=5,
                 nms_padding=2)
FLOP_PER_FORWARD_640 = 298.94e9          # 149.469 GMAC per 640x640 forward


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops": float(p.get("bf16_tflops_sustained",
                                                                      p["bf16_tflops"])),
                "source": "MEASURED_PEAKS.json (sustained bf16, copy HBM)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "source": "B200_PROFILING.md fallback"}


def make_input(batch, size, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((batch, 3, size, size), generator=g)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    # started BEFORE the warm-up (nvidia-smi needs up to a second to come up on an 8-GPU box);
    # only the samples whose timestamp falls inside the timed region are used
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.limit")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.tmp,
                stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0=None, t1=None):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, smax, reasons, watts, wlimit, mask = [], [], set(), [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            ts = self._stamp(parts[0])
            if t0 is not None and ts is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
            # context for a clock below the maximum: board power against its limit and the raw
            # bit mask of active clock-event reasons (0x4 = sw_power_cap)
            try:
                watts.append(float(parts[3]))
                if len(parts) > 9:
                    wlimit.append(float(parts[9]))
            except ValueError:
                pass
            mask.add(parts[4])
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(smax))
            out["samples"] = len(sm)
        if watts:
            out["power_w"] = float(np.median(watts))
        if wlimit:
            out["power_limit_w"] = float(np.median(wlimit))
        if mask:
            out["reason_masks"] = sorted(mask)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------
# CPU implementation of the path.  When the UNMODIFIED reference is importable ($RTPE_REF,
# /root/reference in the build container, or baseline/_ref -- the copy __graft_entry__.build()
# installs, which travels to the GPU box) its own PoseHigherResolutionNet and HeatmapParser are
# timed ("kind": "reference"); otherwise the oracle port ("kind": "port").  The flip-test
# aggregation is upstream HigherHRNet code that the reference imports but does not vendor
# (legacy/valid_ae_avg.py:32-33): both kinds use its restatement oracle/aggregate_ref.py, and
# ``munkres`` (un-vendored, absent) is oracle/munkres_ref.py in both.
# ------------------------------------------------------------------------------------------
class CpuPath:
    def __init__(self, size, mode="auto"):
        import rtpe_b200
        from oracle import group_ref as G
        from oracle import ref_loader
        from oracle.aggregate_ref import aggregate_flip_multiscale_ref
        from oracle.hhrnet_ref import hhrnet_forward_ref
        self.G, self.agg = G, aggregate_flip_multiscale_ref
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.params = G.DecodeParams(**PARSER_KW)
        self.size = size
        self.kind = "port"
        use_ref = mode in ("auto", "reference") and ref_loader.reference_available()
        if use_ref:
            ref_group, ref_model = ref_loader.load_reference()
            torch.manual_seed(0)
            self.net = ref_model.PoseHigherResolutionNet().eval()   # default init, seed 0
            self.parser = ref_group.HeatmapParser(**PARSER_KW)
            self.kind = "reference"
            self.ref_root = ref_loader.REF_ROOT
        else:
            torch.manual_seed(0)
            self.sd = rtpe_b200.PoseHigherResolutionNet().state_dict()   # default init, seed 0
            self.fwd = hhrnet_forward_ref

    def describe(self):
        if self.kind == "reference":
            return ("the UNMODIFIED reference (rtpe.third_party.pose_higher_hrnet + group.HeatmapParser "
                    "from %s) on the CPU; flip aggregation = oracle restatement of the un-vendored "
                    "upstream code, munkres = oracle/munkres_ref.py" % self.ref_root)
        return "oracle port of the reference's per-image loop (no reference tree found)"

    @torch.no_grad()
    def forward(self, x1):
        if self.kind == "reference":
            return [t.float() for t in self.net(x1)]
        return self.fwd(self.sd, x1)

    def decode(self, det, tag):
        """-> (people (P,J,3+T) float32, scores float32[P]) of one image."""
        if self.kind == "reference":
            grouped, scores = self.parser.parse(det, tag, adjust=True, refine=True)
            return np.asarray(grouped[0], np.float32), np.asarray(scores, np.float32)
        people, scores = self.G.parse_image_ref(det.numpy(), tag.numpy(), self.params, True, True)
        return np.asarray(people, np.float32), np.asarray(scores, np.float32)

    @torch.no_grad()
    def one_image(self, x1, keep=False):
        """the reference's per-image loop body: 2 forwards (flip test), aggregation, parse."""
        y = self.forward(x1)
        yf = self.forward(torch.flip(x1, [3]))
        det, tag = self.agg([(1.0, y, yf)], (self.size, self.size))
        people, scores = self.decode(det, tag)
        if keep:
            return {"y": y, "yf": yf, "det": det, "tag": tag, "people": people, "scores": scores}
        return people, scores


def parity_check(cpu, ref, gpu, tol):
    """One image of the timed batch, GPU path against the CPU leg: network outputs and aggregated
    maps within ``tol`` (max|d| / max|ref| per tensor), and the GPU decode of the DEVICE maps
    bit-exact against the oracle decode of the same maps."""
    def rel(got, want):
        return float((got.double().cpu() - want.double()).abs().max() / want.double().abs().max())
    out = {"image": 0, "tolerance": tol, "against": cpu.kind}
    errs = {"y0": rel(gpu["y0"], ref["y"][0]), "y1": rel(gpu["y1"], ref["y"][1]),
            "y0_flip": rel(gpu["y0f"], ref["yf"][0]), "y1_flip": rel(gpu["y1f"], ref["yf"][1]),
            "det": rel(gpu["det"], ref["det"]), "tag": rel(gpu["tag"], ref["tag"])}
    out["max_rel_err"] = errs
    out["float_ok"] = bool(all(v <= tol for v in errs.values()))
    wp, ws = cpu.G.parse_image_ref(gpu["det"].cpu().numpy().copy(), gpu["tag"].cpu().numpy().copy(),
                                   cpu.params, True, True)
    wp = np.asarray(wp, np.float32)
    out["decode_bit_exact"] = bool(gpu["people"].shape == wp.shape and np.array_equal(gpu["people"], wp)
                                   and np.array_equal(gpu["scores"], np.asarray(ws, np.float32)))
    out["people"] = int(wp.shape[0]) if wp.ndim == 3 else 0
    out["people_cpu_leg"] = int(ref["people"].shape[0]) if ref["people"].ndim == 3 else 0
    return out


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cpu = CpuPath(args.size)
    x = make_input(max(1, min(args.batch, args.warmup + args.steps)), args.size)
    for i in range(args.warmup):
        cpu.one_image(x[i % x.shape[0]:i % x.shape[0] + 1])
    t0 = time.perf_counter()
    for i in range(args.steps):
        k = (args.warmup + i) % x.shape[0]
        cpu.one_image(x[k:k + 1])
    dt = time.perf_counter() - t0
    value = args.steps / dt
    sample = ("each step = 1 image of the %dx%d batch: 2 fp32 CPU forwards (flip test) + "
              "aggregation + parse(adjust, refine)" % (args.size, args.size))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cpu.cores, "kind": cpu.kind,
                         "torch_threads": torch.get_num_threads(), "sample": sample,
                         "what": cpu.describe()},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args):
    return {"workload": "HigherHRNet-W48 teacher inference, batch %d synthetic %dx%d per GPU, "
                        "flip test (2 forwards/image), aggregation to %dx%d (T=2), "
                        "HeatmapParser.parse(adjust, refine), random-init weights (seed 0)"
                        % (args.batch, args.size, args.size, args.size, args.size),
            "batch_per_gpu": args.batch, "image_size": args.size, "flip_test": True,
            "precision": args.mode, "chunk": args.chunk,
            "l2": "per-step input (%.0f MB) and activations exceed the 126 MB L2; no flush needed"
                  % (args.batch * 3 * args.size * args.size * 4 / 1e6)}


def conv_bound_split(net, chunk, size, in_dtype, peaks, ms_ops, kinds, flops):
    """Every conv launch against ITS roofline: algorithmic bytes = input + output (+ residual) of the
    layer once (activations are larger than the L2 at the bench's chunk), arithmetic intensity below
    the ridge (tensor peak / HBM peak) = HBM-bound.  -> summary of both groups."""
    plan = net._get_plan(chunk, size, size, net._mode(), net._ref_param().device, in_dtype)
    esz = 2
    ridge = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    grp = {"tensor": [0.0, 0.0, 0.0, 0], "hbm": [0.0, 0.0, 0.0, 0]}     # ms, flop, bytes, launches
    for i, op in enumerate(plan.recorder.ops):
        if op[0] != "conv" or kinds[i] not in (0, 3):
            continue
        d = op[1]
        split = 2 if d.dtype == 2 else 1
        b = d.N * d.Hin * d.Win * d.Cin * esz * split + d.N * d.Hm * d.Wm * d.Cout * esz * split
        if op[5] is not None:
            b += d.N * d.Hm * d.Wm * d.Cout * esz * split
        b += sum(d.N * (d.Hout >> d.add_shift[k]) * (d.Wout >> d.add_shift[k]) * d.Cout * esz
                 for k in range(d.n_add))
        if d.out2_ld:
            b += d.N * d.Hm * d.Wm * d.Cout * esz
        g = grp["hbm" if flops[i] / b < ridge else "tensor"]
        g[0] += ms_ops[i]; g[1] += flops[i]; g[2] += b; g[3] += 1
    out = {"ridge_flop_per_byte": ridge,
           "note": "per launch: algorithmic bytes = layer input + output (+ residual) once"}
    for name, (ms, fl, by, nl) in grp.items():
        out[name + "_bound"] = {
            "launches": nl, "ms": ms, "tflops": fl / max(ms, 1e-9) / 1e9,
            "tensor_frac": fl / max(ms, 1e-9) / 1e9 / peaks["tflops"],
            "gbs": by / max(ms, 1e-9) / 1e6, "hbm_frac": by / max(ms, 1e-9) / 1e6 / peaks["hbm_gbs"]}
    return out


def config5_leg(peaks):
    """BASELINE.json configs[4]: decode-only stress, batch 1024, 17 x 320 x 320, T = 1, K = 30
    (SURVEY.md 8d generator); device time per stage, algorithmic bytes against the HBM peak."""
    import rtpe_b200
    n, sz = 1024, 320
    parts = [rtpe_b200.synth_decode_batch(128, height=sz, width=sz, tag_dims=1, max_people=30,
                                          seed=1234, device="cuda", first_index=i)
             for i in range(0, n, 128)]
    det = torch.cat([q[0] for q in parts])
    tag = torch.cat([q[1] for q in parts])
    del parts
    parser = rtpe_b200.HeatmapParser(**PARSER_KW)
    for _ in range(2):
        parser.decode_device(det, tag, full_capacity=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    torch.cuda._sleep(8_000_000)
    ev[0].record()
    val_k, ind_k, _, tag_k = parser.top_k_device(det, tag)
    ev[1].record()
    ans, count, pmax = parser.match_device(val_k, ind_k, tag_k, sz, pmax=parser._pmax_full())
    ev[2].record()
    parser.adjust_device(ans, count, det)
    ev[3].record()
    parser.refine_device(det, tag, ans, count)
    ev[4].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    b_topk, b_ref = det.numel() * 4, (det.numel() + tag.numel()) * 4
    total = sum(t)
    pk = peaks["hbm_gbs"]
    return {"workload": "configs[4]: decode-only stress, batch %d, 17 x %d x %d, T=1, K=30" % (n, sz, sz),
            "value": n / total * 1e3, "unit": "images/s", "people_per_image": float(count.float().mean()),
            "ms": {"top_k": t[0], "match": t[1], "adjust": t[2], "refine": t[3], "parse_total": total},
            "top_k_frac": b_topk / t[0] / 1e6 / pk, "refine_frac": b_ref / t[3] / 1e6 / pk,
            "parse_frac": (b_topk + b_ref) / total / 1e6 / pk, "hbm_peak_gbs": pk}


def conv_rooflines(net, chunk, size, in_dtype, peaks, mode):
    """(roofline of conv_halo_kernel, roofline of conv_umma_kernel, per-op ms) of one plan replay:
    algorithmic FLOP (2 * MAC of the reference layer) / CUDA-event time of the launches."""
    net.plan_profile(chunk, size, size, in_dtype)
    ms_ops, kinds, flops = net.plan_profile(chunk, size, size, in_dtype)
    conv_ms = sum(m for m, k in zip(ms_ops, kinds) if k in (0, 1, 3))

    def one(kset, name):
        ms = sum(m for m, k in zip(ms_ops, kinds) if k in kset)
        fl = sum(f for f, k in zip(flops, kinds) if k in kset)
        nl = sum(1 for k in kinds if k in kset)
        ach = fl / max(ms, 1e-9) / 1e9
        return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peaks["tflops"],
                "unit": "TFLOP/s", "frac": ach / peaks["tflops"], "traffic": None, "launches": nl,
                "avg_launch_ms": ms / max(nl, 1), "flop_per_launch_avg": fl / max(nl, 1),
                "share_of_forward": ms / max(sum(ms_ops), 1e-9), "peak_source": peaks["source"],
                "chunk": chunk}

    # dominant kernel: the 3x3/s1 tcgen05 halo kernel (kind 3); the per-tap tcgen05 kernel (1x1,
    # stride 2, deconv phases, stem) is reported beside it
    roofline = one((3,), "conv_halo_kernel (tcgen05 implicit GEMM, 3x3 s1)")
    other = one((0,), "conv_umma_kernel (tcgen05 implicit GEMM, 1x1 / s2 / deconv / 3x3 384ch on 20x20 maps)")
    if mode == "fp32":
        if roofline["launches"] == 0:          # BRTPE engine override: CUDA-core float32 path
            roofline = one((1,), "conv_ffma_kernel (fp32 FFMA implicit GEMM)")
            roofline["note"] = "fp32 CUDA-core path; frac is against the bf16 tensor peak"
        else:
            note = ("fp32 mode = split (hi, lo) bf16 activations, THREE bf16 tcgen05 MMAs per product "
                    "(BRTPE_DT_BF16X2); achieved = algorithmic FLOP of the layer / time, so the "
                    "tensor pipe does 3x this work: frac * 3 is the pipe's share of the bf16 peak")
            roofline["note"] = note
            roofline["pipe_frac"] = 3 * roofline["frac"]
            other["pipe_frac"] = 3 * other["frac"]
    roofline["conv_share_of_forward"] = conv_ms / max(sum(ms_ops), 1e-9)
    try:
        roofline["by_bound"] = conv_bound_split(net, chunk, size, in_dtype, peaks, ms_ops, kinds, flops)
    except Exception as exc:                     # reporting only: never fail the bench on it
        roofline["by_bound"] = {"error": repr(exc)}
    return roofline, other, ms_ops


def config3_leg(args, rank, world, dev, pipe, net, parser):
    """BASELINE.json configs[2] (legacy/valid_ae_avg.py:159-205): synthetic batch 256 sharded over
    the ranks (inference.shard_range), scales 2.0 / 1.0 / 0.5 (inputs 1280^2 / 640^2 / 320^2,
    transforms.py:155-176 with min_scale 0.5), flip test, projection to 640 x 640, T = 2, AE grouping.
    STRONG scaling: the job is the same 256 images whatever N.  Device-resident, timed with CUDA
    events, max over ranks; the per-rank results are all-gathered once per batch."""
    import torch.distributed as dist
    from rtpe_b200 import inference
    total, sub, size = 256, 32, args.size
    lo, hi = inference.shard_range(total, rank, world)
    shard = hi - lo
    g = torch.Generator().manual_seed(100 + rank)
    xs = [(sc, torch.randn(sub, 3, int(size * sc), int(size * sc), generator=g).to(dev))
          for sc in (2.0, 1.0, 0.5)]
    gatherer = inference.ResultGatherer()
    pcap = parser.person_capacity

    def one_batch():
        outs = []
        for s0 in range(0, shard, sub):                 # the shard in passes of `sub` images
            det, tag = pipe.forward_aggregate_multiscale(xs, (size, size))
            ans, count, scores = parser.decode_device(det, tag, True, True, full_capacity=True)
            outs.append(inference.pack_results(*inference.pad_results(ans, count, scores, pcap)))
        payload = torch.cat(outs, 0) if len(outs) > 1 else outs[0]
        out, ev = gatherer.gather(payload)
        torch.cuda.current_stream(dev).wait_event(ev)
        return out

    one_batch()
    dist.barrier()
    torch.cuda.synchronize()
    reps = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = one_batch()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    people = float(out[:, 0].contiguous().view(torch.int32).float().mean().item())
    return {"workload": "configs[2]: batch 256 synthetic 640x640 sharded %d per GPU, scales 2.0/1.0/0.5 "
                        "+ flip test, projected to 640x640 (T=2), decode; %d images per pass" % (shard, sub),
            "scaling": "strong", "n_gpus": world, "images": total, "ms_per_batch": ms,
            "value": total / (ms * 1e-3), "unit": "images/s", "per_gpu": total / (ms * 1e-3) / world,
            "forward_tflops_effective": 2 * (4 + 1 + 0.25) * FLOP_PER_FORWARD_640 * total / (ms * 1e-3) / 1e12,
            "people_per_image": people, "gathered_rows": int(out.shape[0])}


def fp32_leg(args, dev, x_dev, ref, peaks):
    """configs[1] names bf16 AND fp32: the same step with float32 parameters (split-bf16 tcgen05
    path, <= 1e-4), measured next to the bf16 headline and checked against the CPU leg."""
    import rtpe_b200
    from rtpe_b200 import inference
    torch.manual_seed(0)
    model = rtpe_b200.get_hrnet_w48_teacher(None, half=False).to(dev)
    model.chunk_size = args.chunk
    model.freeze()
    parser = rtpe_b200.HeatmapParser(**PARSER_KW)
    pipe = inference.TeacherPipeline(model, parser, flip_test=True)
    for _ in range(2):
        pipe.run_device(x_dev, True, True)
    torch.cuda.synchronize()
    steps = max(2, min(args.steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        pipe.run_device(x_dev, True, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    roofline, other, _ = conv_rooflines(model, min(args.chunk, 2 * args.batch), args.size,
                                        torch.float32, peaks, "fp32")
    out = {"value": x_dev.shape[0] / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "steps": steps,
           "dtype": "fp32 (split bf16x2 on tcgen05)", "roofline": roofline,
           "roofline_other_convs": other, "scope": "device-resident, 1 GPU, same step as `value`"}
    if ref is not None:
        n = x_dev.shape[0]
        with torch.no_grad():
            y0, y1 = pipe._forward_flip(x_dev)
            errs = []
            for got, want in ((y0[0:1], ref["y"][0]), (y1[0:1], ref["y"][1]),
                              (y0[n:n + 1], ref["yf"][0]), (y1[n:n + 1], ref["yf"][1])):
                errs.append(float((got.double().cpu() - want.double()).abs().max() /
                                  want.double().abs().max()))
        out["max_rel_err_vs_cpu_leg"] = max(errs)
        out["parity_checked"] = bool(max(errs) <= 1e-4)
    return out


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import rtpe_b200
    from rtpe_b200 import inference

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cores = inference.pin_to_gpu_numa(local_rank)   # host threads next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    if args.mode == "bf16":
        model = rtpe_b200.get_hrnet_w48_teacher(None).to(dev)
        net = model[1]
    else:
        model = rtpe_b200.get_hrnet_w48_teacher(None, half=False).to(dev)
        net = model
    net.chunk_size = args.chunk
    net.freeze()
    parser = rtpe_b200.HeatmapParser(**PARSER_KW)
    pipe = inference.TeacherPipeline(model, parser, flip_test=True)

    # every rank gets its own shard of the (weak-scaled) global batch
    x_host = make_input(args.batch, args.size, seed=1 + rank).pin_memory()
    x_dev = x_host.to(dev)
    pcap = parser.person_capacity

    # The only collective: ONE packed all-gather of the padded per-image results per step, on a side
    # stream, so that it overlaps the next step's network; the gathered tensor stays on the device,
    # every rank copies only its own shard to the host (d2h bytes per rank do not grow with N).
    gatherer = inference.ResultGatherer() if world > 1 else None
    pending = {"ev": None, "slot": 0}

    def gather_async(ans, count, scores):
        if gatherer is None:
            return
        if pending["ev"] is not None:          # the previous step's gather (other slot) must be done
            torch.cuda.current_stream(dev).wait_event(pending["ev"])
        _, pending["ev"] = gatherer.gather(inference.pack_results(ans, count, scores), pending["slot"])
        pending["slot"] ^= 1

    def gather_drain():
        if pending["ev"] is not None:
            torch.cuda.current_stream(dev).wait_event(pending["ev"])
            pending["ev"] = None

    def step_device():
        ans, count, scores = pipe.run_device(x_dev, True, True)
        ans, count, scores = inference.pad_results(ans, count, scores, pcap)
        gather_async(ans, count, scores)
        return ans, count, scores

    host_out = {}

    def finish_e2e(ans, count, scores):
        ans, count, scores = inference.pad_results(ans, count, scores, pcap)
        gather_async(ans, count, scores)
        for k, t in (("ans", ans), ("count", count), ("scores", scores)):
            if k not in host_out:
                host_out[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            host_out[k].copy_(t, non_blocking=True)

    def run_e2e(nsteps):
        """nsteps batches through the host-facing pipelined call: every step copies its input
        from pinned host memory (overlapped with the previous step's kernels) and copies its
        results back to pinned host memory."""
        for ans, count, scores in pipe.run_stream((x_host for _ in range(nsteps)), True, True):
            finish_e2e(ans, count, scores)
        gather_drain()

    def timed(fn, steps, warmup, sample_clocks):
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        gather_drain()                     # the last step's gather belongs to the timed region
        e1.record()
        torch.cuda.synchronize()
        t_wall1 = time.time()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks

    ms_dev, clocks = timed(step_device, args.steps, args.warmup, True)
    run_e2e(max(2, args.warmup))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_t.item())
    ans, count, scores = step_device()
    gather_drain()
    torch.cuda.synchronize()
    people_mean = float(count.float().mean().item())

    total_images = args.batch * world * args.steps
    value = total_images / (ms_dev * 1e-3)
    e2e_value = total_images / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv), measured live with
    # CUDA events around every launch of one plan replay, on the launching stream
    peaks = load_peaks()
    in_dtype = torch.float16 if args.mode == "bf16" else torch.float32
    chunk = min(args.chunk, 2 * args.batch)
    roofline, roofline_other, ms_ops = conv_rooflines(net, chunk, args.size, in_dtype, peaks,
                                                      args.mode)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath) and args.mode == "bf16":
        with open(tpath) as f:
            t = json.load(f)
        roofline["traffic"] = t.get("dram_bytes_per_launch")
        roofline["traffic_note"] = t.get("note")

    # ---- decode kernels against the HBM roofline (secondary)
    det, tag = pipe.forward_aggregate(x_dev)            # the step's own maps, whole batch
    nd, j, h, w = det.shape
    t = tag.shape[4]
    for _ in range(2):
        parser.decode_device(det, tag)
    torch.cuda.synchronize()
    # aggregation kernel alone (network outputs -> det/tag), on the step's own outputs
    nimg = x_dev.shape[0]
    both = torch.cat((x_dev, torch.flip(x_dev, [3])), 0)
    y0, y1 = model(both)
    del both
    agg_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    inference.aggregate_scale(y0[:nimg], y1[:nimg], y0[nimg:], y1[nimg:], (args.size, args.size), 17)
    # the GPU spins ~2 ms first so that the host is ahead of the device when the events are
    # recorded: the intervals are device time of the kernels, not host launch latency
    torch.cuda._sleep(4_000_000)
    agg_ev[0].record()
    inference.aggregate_scale(y0[:nimg], y1[:nimg], y0[nimg:], y1[nimg:], (args.size, args.size), 17)
    agg_ev[1].record()
    torch.cuda.synchronize()
    agg_ms = agg_ev[0].elapsed_time(agg_ev[1])
    agg_bytes = (det.numel() + tag.numel() + y0.numel() + y1.numel()) * 4
    del y0, y1
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    torch.cuda._sleep(4_000_000)
    ev[0].record()
    val_k, ind_k, _, tag_k = parser.top_k_device(det, tag)
    ev[1].record()
    a2, c2, _ = parser.match_device(val_k, ind_k, tag_k, w)      # ends with a host sync (overflow flag)
    ev[2].record()
    torch.cuda._sleep(4_000_000)
    ev[5].record()
    parser.adjust_device(a2, c2, det)
    ev[3].record()
    parser.refine_device(det, tag, a2, c2)
    ev[4].record()
    torch.cuda.synchronize()
    b_topk = nd * j * h * w * 4
    b_refine = nd * j * h * w * 4 * (1 + t)
    roofline_decode = {
        "bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"],
        "top_k": {"achieved": b_topk / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9,
                  "ms": ev[0].elapsed_time(ev[1])},
        "refine": {"achieved": b_refine / (ev[3].elapsed_time(ev[4]) * 1e-3) / 1e9,
                   "ms": ev[3].elapsed_time(ev[4])},
        "aggregate": {"achieved": agg_bytes / (agg_ms * 1e-3) / 1e9, "ms": agg_ms,
                      "frac": agg_bytes / (agg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "bytes": "reads every network output once + writes det and tag once"},
        "match_ms": ev[1].elapsed_time(ev[2]), "images": nd,
        "bytes_per_image": 4 * j * h * w * (2 + t)}
    roofline_decode["top_k"]["frac"] = roofline_decode["top_k"]["achieved"] / peaks["hbm_gbs"]
    roofline_decode["refine"]["frac"] = roofline_decode["refine"]["achieved"] / peaks["hbm_gbs"]

    plan_ops = len(ms_ops)
    n_chunks = -(-2 * args.batch // args.chunk)
    gpu_launches = (n_chunks * plan_ops + 1 + 9) * args.steps

    cpu_baseline, parity, ref = None, None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = CpuPath(args.size)
        xs = x_host[:1].clone()                       # image 0 of the timed batch
        t0 = time.perf_counter()
        ref = cpu.one_image(xs, keep=True)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": 1.0 / dt, "unit": "images/s", "cores": cpu.cores, "kind": cpu.kind,
                        "torch_threads": torch.get_num_threads(), "seconds": dt,
                        "sample": "image 0 of the timed batch: 2 fp32 CPU forwards (flip test) + "
                                  "aggregation + parse(adjust, refine)", "what": cpu.describe()}
        # the same image through the GPU path, exactly as the timed step runs it (whole batch)
        nimg = x_dev.shape[0]
        with torch.no_grad():
            gy0, gy1 = pipe._forward_flip(x_dev)
            gpu = {"y0": gy0[0:1].clone(), "y1": gy1[0:1].clone(), "y0f": gy0[nimg:nimg + 1].clone(),
                   "y1f": gy1[nimg:nimg + 1].clone()}
            gdet, gtag = pipe.forward_aggregate(x_dev)
            gans, gcount, gscores = parser.decode_device(gdet, gtag, True, True)
        c0 = int(gcount[0])
        gpu.update(det=gdet[0:1], tag=gtag[0:1], people=gans[0, :c0].cpu().numpy(),
                   scores=gscores[0, :c0].cpu().numpy())
        parity = parity_check(cpu, ref, gpu, 2e-2 if args.mode == "bf16" else 1e-4)
    # Optional legs (extra keys of the line).  The headline numbers above are complete at this point: a
    # failure in a leg is reported in its key instead of costing the whole line (a CUDA launch failure
    # poisons the context, so the process then leaves right after printing).
    leg_failed = False
    fp32 = None
    if rank == 0 and world == 1 and args.mode == "bf16" and not args.no_fp32:
        try:
            fp32 = fp32_leg(args, dev, x_dev, ref if parity else None, peaks)
        except Exception as exc:                 # noqa: BLE001
            fp32 = {"error": repr(exc)[:300]}
            leg_failed = True

    config5 = None
    if rank == 0 and world == 1 and not args.no_config5 and not leg_failed:
        try:
            config5 = config5_leg(peaks)
        except Exception as exc:                 # noqa: BLE001
            config5 = {"error": repr(exc)[:300]}
            leg_failed = True
    config3 = None
    if world > 1 and args.mode == "bf16" and not args.no_config3:
        config3 = config3_leg(args, rank, world, dev, pipe, net, parser)

    if rank == 0:
        h2d = x_host.numel() * x_host.element_size()
        d2h = sum(v.numel() * v.element_size() for v in host_out.values())
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.mode, "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline,
            "roofline_other_convs": roofline_other,
            "roofline_decode": roofline_decode, "cpu_baseline": cpu_baseline,
            "parity_checked": bool(parity and parity["float_ok"] and parity["decode_bit_exact"]),
            "parity": parity, "fp32": fp32, "config3": config3, "config5_decode": config5,
            "host": {"numa_cores_bound": numa_cores, "cpu_count": os.cpu_count()},
            "forward_tflops_effective": 2 * args.batch * world * args.steps *
            FLOP_PER_FORWARD_640 * (args.size / 640.0) ** 2 / (ms_dev * 1e-3) / 1e12,
            "people_per_image": people_mean,
        }
        emit(line)
        if leg_failed:
            sys.stdout.flush()
            sys.stderr.write("bench.py: an optional leg failed (see its key in the line); leaving without teardown\n")
            sys.stderr.flush()
            os._exit(0)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--chunk", type=int, default=64,
                    help="forwards per plan replay (64 = the whole flip-test batch in one CUDA graph)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-mode leg of the bf16 line")
    ap.add_argument("--no-config5", action="store_true",
                    help="skip the decode-only stress leg (BASELINE configs[4]) of the 1-GPU line")
    ap.add_argument("--no-config3", action="store_true",
                    help="skip the multi-scale batch-256 leg (BASELINE configs[2]) of multi-GPU runs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 2
        args.warmup = min(args.warmup, 1)
        run_reference_arm(args, rank)
        return
    if args.steps is None:
        args.steps = 10
    if world != args.gpus and world > 1:
        print("warning: WORLD_SIZE %d != --gpus %d" % (world, args.gpus), file=sys.stderr)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
