"""Config 4 of BASELINE.json: AttentionStudent (inplanes 48, 17 heat-maps + 1 AE map) forward + the same
decode (tag_per_joint=False, rtpe/engine.py:45-49), batch 64 synthetic 512 x 512, fp32 (the reference's
half_precision=False) and bf16."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rtpe_b200  # noqa: E402


def run(half, batch, size, reps):
    torch.manual_seed(0)
    net = rtpe_b200.AttentionStudent(None, "cuda", inplanes=48, num_heatmaps=17, ae_dims=1,
                                     half_precision=half).eval()
    net.chunk_size = batch
    net.freeze()
    parser = rtpe_b200.HeatmapParser(17, 30, 0.1, 1.0, True, False, tag_per_joint=False, nms_ksize=5,
                                     nms_padding=2)
    x = torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(1)).cuda()

    def step():
        with torch.no_grad():
            att, det = net(x)
        return parser.decode_device(det[:, :17].contiguous(), det[:, 17:].unsqueeze(-1).contiguous())

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(reps):
        with torch.no_grad():
            att, det = net(x)
    ev[1].record()
    for _ in range(reps):
        ans, count, scores = step()
    ev[2].record()
    torch.cuda.synchronize()
    fwd = ev[0].elapsed_time(ev[1]) / reps
    full = ev[1].elapsed_time(ev[2]) / reps
    return {"mode": "bf16" if half else "fp32", "forward_ms": fwd, "forward_decode_ms": full,
            "images_per_s": batch / full * 1e3, "forward_tflops": 2 * 13.9e9 * batch / fwd / 1e9,
            "people_per_image": float(count.float().mean())}


def main():
    batch, size = int(os.environ.get("BATCH", "64")), int(os.environ.get("SIZE", "512"))
    out = {"workload": "config 4: AttentionStudent batch %d, %dx%d, forward + decode at %dx%d" % (
        batch, size, size, size // 4, size // 4)}
    out["bf16"] = run(True, batch, size, 5)
    out["fp32"] = run(False, batch, size, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
