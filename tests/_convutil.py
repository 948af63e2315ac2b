"""Helpers for the conv parity tests: drive brtpe_conv_run for a standard Conv2d and build
the torch reference of the same op (the product never uses torch convs)."""
import ctypes as C

import torch
import torch.nn.functional as F

from rtpe_b200 import _lib as L

TAPS3 = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]


def make_desc(dtype, engine, n, h, w, cin, cout, k, stride, relu, in_ld=None, in_coff=0,
              out_ld=None, out_coff=0, cout_store=None, res_ld=0):
    d = L.ConvDesc()
    taps = TAPS3 if k == 3 else [(0, 0)]
    ho, wo = h // stride, w // stride
    cout_store = cout if cout_store is None else cout_store
    d.dtype, d.engine = dtype, engine
    d.N, d.Hin, d.Win = n, h, w
    d.Cin, d.in_ld, d.in_coff = cin, (in_ld or cin), in_coff
    d.Hm, d.Wm, d.in_stride = ho, wo, stride
    d.ntaps = len(taps)
    for i, (dy, dx) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i] = dy, dx
    d.Hout, d.Wout, d.out_scale, d.out_oy, d.out_ox = ho, wo, 1, 0, 0
    d.Cout = cout
    d.out_ld = out_ld or ((cout_store + 15) // 16 * 16)
    d.out_coff = out_coff
    d.res_ld, d.res_coff = res_ld, 0
    d.relu = int(relu)
    d.Cout_store = cout_store
    return d, taps


def pack_weights(lib, w, taps, k, desc, engine_used, bf16_round):
    """w (Cout,Cin,k,k) f32 -> packed tensor for engine_used."""
    wt = torch.stack([w[:, :, dy + k // 2, dx + k // 2] for dy, dx in taps], 0)  # (T,Cout,Cin)
    if engine_used in (L.ENGINE_UMMA, L.ENGINE_UMMA_HALO):
        cp, op = C.c_int(0), C.c_int(0)
        lib.brtpe_umma_weight_dims(desc.Cin, desc.Cout_store, C.byref(cp), C.byref(op))
        packed = torch.zeros((len(taps), op.value, cp.value), dtype=torch.bfloat16, device=w.device)
        packed[:, :w.shape[0], :w.shape[1]] = wt.to(torch.bfloat16)
        return packed
    if bf16_round:
        wt = wt.to(torch.bfloat16).float()
    return wt.permute(0, 2, 1).contiguous()


def run_conv(engine, mode, n, h, w, cin, cout, k, stride, relu, use_res, seed=0, device="cuda"):
    """-> (out NCHW f32 from the library, reference NCHW f32, engine used)."""
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    bf = mode == "bf16"
    tdt = torch.bfloat16 if bf else torch.float32
    x = torch.randn((n, cin, h, w), generator=g).to(device)
    wgt = (torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5).to(device)
    bias = torch.randn((cout,), generator=g).to(device) * 0.1
    ho, wo = h // stride, w // stride
    res = torch.randn((n, cout, ho, wo), generator=g).to(device) if use_res else None
    d, taps = make_desc(L.DT_BF16 if bf else L.DT_F32, engine, n, h, w, cin, cout, k, stride,
                        relu, res_ld=(cout if use_res else 0))
    eng = lib.brtpe_conv_select_engine(C.byref(d))
    assert eng in (L.ENGINE_FFMA, L.ENGINE_UMMA, L.ENGINE_UMMA_HALO), lib.brtpe_last_error()
    xin = x.permute(0, 2, 3, 1).contiguous().to(tdt)
    rin = res.permute(0, 2, 3, 1).contiguous().to(tdt) if use_res else None
    packed = pack_weights(lib, wgt, taps, k, d, eng, bf)
    out = torch.full((n, ho, wo, d.out_ld), float("nan"), dtype=tdt, device=device)
    L.check(lib.brtpe_conv_run(C.byref(d), L.ptr(xin), L.ptr(packed), L.ptr(bias), L.ptr(rin),
                               L.ptr(out), L.stream_ptr()), "brtpe_conv_run")
    torch.cuda.synchronize()
    xr = xin.float().permute(0, 3, 1, 2)
    wr = wgt.to(torch.bfloat16).float() if bf else wgt
    ref = F.conv2d(xr.double(), wr.double(), bias.double(), stride=stride, padding=k // 2)
    if use_res:
        ref = ref + rin.double().permute(0, 3, 1, 2)
    if relu:
        ref = F.relu(ref)
    got = out[..., :cout].float().permute(0, 3, 1, 2)
    return got, ref.float(), eng


def split_hi_lo(t):
    """float32 -> (hi, lo) bf16 pair of BRTPE_DT_BF16X2: hi = bf16(v), lo = bf16(v - hi)."""
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


def run_conv_split(engine, n, h, w, cin, cout, k, stride, relu, use_res, seed=0, device="cuda"):
    """fp32 mode of the tcgen05 engines (BRTPE_DT_BF16X2): split activations, weights packed by
    brtpe_prepack_weights(split=1).  -> (out NCHW f32 = hi + lo, float64 reference, engine)."""
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((n, cin, h, w), generator=g).to(device)
    wgt = (torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5).to(device)
    bias = torch.randn((cout,), generator=g).to(device) * 0.1
    ho, wo = h // stride, w // stride
    res = torch.randn((n, cout, ho, wo), generator=g).to(device) if use_res else None
    d, taps = make_desc(L.DT_BF16X2, engine, n, h, w, cin, cout, k, stride, relu, in_ld=2 * cin,
                        out_ld=2 * cout, res_ld=(2 * cout if use_res else 0))
    eng = lib.brtpe_conv_select_engine(C.byref(d))
    assert eng in (L.ENGINE_UMMA, L.ENGINE_UMMA_HALO), lib.brtpe_last_error()

    def to_split_nhwc(t):
        hi, lo = split_hi_lo(t.permute(0, 2, 3, 1).contiguous())
        return torch.cat((hi, lo), dim=3).contiguous()
    xin = to_split_nhwc(x)
    rin = to_split_nhwc(res) if use_res else None
    cp, op = C.c_int(0), C.c_int(0)
    lib.brtpe_umma_weight_dims(cin, cout, C.byref(cp), C.byref(op))
    pd = L.PrepackDesc()
    pd.w_dtype, pd.transposed = L.WT_F32, 0
    pd.Cout, pd.Cin, pd.KH, pd.KW = cout, cin, k, k
    pd.ntaps = len(taps)
    for i, (dy, dx) in enumerate(taps):
        pd.tap_kh[i], pd.tap_kw[i] = dy + k // 2, dx + k // 2
    pd.Cin_store, pd.layout, pd.split = cin, L.PACK_KMAJOR_BF16, 1
    pd.cin_pad, pd.cout_pad = 3 * cp.value, op.value
    packed = torch.empty((len(taps), op.value, 3 * cp.value), dtype=torch.bfloat16, device=device)
    L.check(lib.brtpe_prepack_weights(C.byref(pd), L.ptr(wgt), None, None, None, None, None, None,
                                      L.ptr(packed), None, 0, L.stream_ptr()), "brtpe_prepack_weights")
    out = torch.full((n, ho, wo, 2 * cout), float("nan"), dtype=torch.bfloat16, device=device)
    L.check(lib.brtpe_conv_run(C.byref(d), L.ptr(xin), L.ptr(packed), L.ptr(bias), L.ptr(rin),
                               L.ptr(out), L.stream_ptr()), "brtpe_conv_run")
    torch.cuda.synchronize()
    xr = (xin[..., :cin].double() + xin[..., cin:].double()).permute(0, 3, 1, 2)
    ref = F.conv2d(xr, wgt.double(), bias.double(), stride=stride, padding=k // 2)
    if use_res:
        ref = ref + (rin[..., :cout].double() + rin[..., cout:].double()).permute(0, 3, 1, 2)
    if relu:
        ref = F.relu(ref)
    got = (out[..., :cout].float() + out[..., cout:].float()).permute(0, 3, 1, 2)
    return got, ref.float(), eng


def run_conv_fused(engine, n, h, w, cin, cout, k, stride, relu, use_res, shifts, dual, seed=0,
                   device="cuda"):
    """HRNet fuse-add in the epilogue (brtpe_conv_run_fused): addends at resolution >> shift, read
    with nearest-neighbour upsampling.  -> (out, out2 or None, ref_out, ref_out2 or None, engine)."""
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((n, cin, h, w), generator=g).to(device)
    wgt = (torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5).to(device)
    bias = torch.randn((cout,), generator=g).to(device) * 0.1
    ho, wo = h // stride, w // stride
    res = torch.randn((n, cout, ho, wo), generator=g).to(device) if use_res else None
    adds = [torch.randn((n, cout, ho >> s, wo >> s), generator=g).to(device) for s in shifts]
    d, taps = make_desc(L.DT_BF16, engine, n, h, w, cin, cout, k, stride, relu,
                        res_ld=(cout if use_res else 0))
    d.n_add = len(shifts)
    for i, s in enumerate(shifts):
        d.add_ld[i], d.add_shift[i] = cout, s
    d.out2_ld = cout if dual else 0
    eng = lib.brtpe_conv_select_engine(C.byref(d))
    assert eng in (L.ENGINE_UMMA, L.ENGINE_UMMA_HALO), lib.brtpe_last_error()
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    xin, rin = nhwc(x), (nhwc(res) if use_res else None)
    ain = [nhwc(a) for a in adds]
    packed = pack_weights(lib, wgt, taps, k, d, eng, True)
    out = torch.full((n, ho, wo, d.out_ld), float("nan"), dtype=torch.bfloat16, device=device)
    out2 = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device=device) if dual else None
    ap = (C.c_void_p * max(len(ain), 1))(*[a.data_ptr() for a in ain])
    L.check(lib.brtpe_conv_run_fused(C.byref(d), L.ptr(xin), L.ptr(packed), L.ptr(bias), L.ptr(rin),
                                     L.ptr(out), ap, L.ptr(out2), L.stream_ptr()), "brtpe_conv_run_fused")
    torch.cuda.synchronize()
    ref = F.conv2d(xin.double().permute(0, 3, 1, 2), wgt.to(torch.bfloat16).double(), bias.double(),
                   stride=stride, padding=k // 2)
    if use_res:
        ref = ref + rin.double().permute(0, 3, 1, 2)
    tot = 0
    for a, s in zip(ain, shifts):
        up = a.double().permute(0, 3, 1, 2)
        if s:
            up = F.interpolate(up, scale_factor=2 ** s, mode="nearest")
        tot = tot + up
    got = out[..., :cout].float().permute(0, 3, 1, 2)
    if dual:
        r1 = F.relu(ref) if relu else ref
        r2 = F.relu(r1 + tot)
        return got, out2.float().permute(0, 3, 1, 2), r1.float(), r2.float(), eng
    r1 = ref + tot
    r1 = F.relu(r1) if relu else r1
    return got, None, r1.float(), None, eng


def run_chain(n, h, w, c, seed=0, device="cuda", env_g=None):
    """BasicBlock body (conv1 + relu -> conv2 + residual + relu, bf16, halo engine) as ONE chained launch
    (brtpe_conv_chain_run) and as two brtpe_conv_run calls -> (mid, out) of both, NHWC bf16."""
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((n, h, w, c), generator=g).to(device).to(torch.bfloat16)
    ws = [(torch.randn((c, c, 3, 3), generator=g) / (c * 9) ** 0.5).to(device) for _ in range(2)]
    bs = [(torch.randn((c,), generator=g) * 0.1).to(device) for _ in range(2)]
    d0, taps = make_desc(L.DT_BF16, L.ENGINE_UMMA_HALO, n, h, w, c, c, 3, 1, True)
    d1, _ = make_desc(L.DT_BF16, L.ENGINE_UMMA_HALO, n, h, w, c, c, 3, 1, True, res_ld=c)
    pk = [pack_weights(lib, ws[i], taps, 3, d0, L.ENGINE_UMMA_HALO, True) for i in range(2)]
    res = {}
    for name in ("two", "chain"):
        mid = torch.full((n, h, w, d0.out_ld), float("nan"), dtype=torch.bfloat16, device=device)
        out = torch.full((n, h, w, d1.out_ld), float("nan"), dtype=torch.bfloat16, device=device)
        if name == "two":
            L.check(lib.brtpe_conv_run(C.byref(d0), L.ptr(x), L.ptr(pk[0]), L.ptr(bs[0]), None,
                                       L.ptr(mid), L.stream_ptr()), "brtpe_conv_run")
            L.check(lib.brtpe_conv_run(C.byref(d1), L.ptr(mid), L.ptr(pk[1]), L.ptr(bs[1]), L.ptr(x),
                                       L.ptr(out), L.stream_ptr()), "brtpe_conv_run")
        else:
            L.check(lib.brtpe_conv_chain_run(C.byref(d0), L.ptr(x), L.ptr(pk[0]), L.ptr(bs[0]), L.ptr(mid),
                                             C.byref(d1), L.ptr(pk[1]), L.ptr(bs[1]), L.ptr(x),
                                             L.ptr(out), L.stream_ptr()), "brtpe_conv_chain_run")
        torch.cuda.synchronize()
        res[name] = (mid, out)
    return res
