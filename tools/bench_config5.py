"""Config 5 of BASELINE.json: decode-only stress -- synthetic 17-joint heat-maps + tags at 320x320,
max_num_people = 30, batch 1024 (SURVEY 8d generator).  Device time of every decode stage and the
parse() roofline (algorithmic bytes 4*J*H*W*(2+T) per image against the measured copy bandwidth)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rtpe_b200  # noqa: E402


def main():
    n = int(os.environ.get("BATCH", "1024"))
    s = int(os.environ.get("SIZE", "320"))
    parts = [rtpe_b200.synth_decode_batch(128, height=s, width=s, tag_dims=1, max_people=30, seed=1234,
                                          device="cuda", first_index=i) for i in range(0, n, 128)]
    det = torch.cat([p[0] for p in parts])
    tag = torch.cat([p[1] for p in parts])
    del parts
    parser = rtpe_b200.HeatmapParser(17, 30, 0.1, 1.0, True, False, nms_ksize=5, nms_padding=2)
    for _ in range(2):
        parser.decode_device(det, tag)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    torch.cuda._sleep(8_000_000)
    ev[0].record()
    val_k, ind_k, _, tag_k = parser.top_k_device(det, tag)
    ev[1].record()
    ans, count, _ = parser.match_device(val_k, ind_k, tag_k, s)
    ev[2].record()
    torch.cuda._sleep(8_000_000)
    ev[3].record()
    parser.adjust_device(ans, count, det)
    ev[4].record()
    parser.refine_device(det, tag, ans, count)
    ev[5].record()
    torch.cuda.synchronize()
    peak = 6456.5
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)["hbm_gbs"]
    except OSError:
        pass
    t_topk, t_match = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    t_adj, t_ref = ev[3].elapsed_time(ev[4]), ev[4].elapsed_time(ev[5])
    b_topk, b_ref = det.numel() * 4, (det.numel() + tag.numel()) * 4
    total = t_topk + t_match + t_adj + t_ref
    print(json.dumps({
        "workload": "config 5: decode-only, batch %d, 17 x %d x %d, T=1, K=30" % (n, s, s),
        "people_per_image": float(count.float().mean()),
        "ms": {"top_k": t_topk, "match": t_match, "adjust": t_adj, "refine": t_ref, "parse_total": total},
        "images_per_s": n / total * 1e3,
        "hbm_peak_gbs": peak,
        "top_k_gbs": b_topk / t_topk / 1e6, "top_k_frac": b_topk / t_topk / 1e6 / peak,
        "refine_gbs": b_ref / t_ref / 1e6, "refine_frac": b_ref / t_ref / 1e6 / peak,
        "parse_gbs": (b_topk + b_ref) / total / 1e6, "parse_frac": (b_topk + b_ref) / total / 1e6 / peak}))


if __name__ == "__main__":
    main()
