"""Per-pass timing of the implicit-GEMM convolutions on the hot-path shapes, next to cuDNN (channels_last, TF32, benchmark mode).
CUDA events, L2 flushed between repetitions.  `python scripts/conv_bench.py [--quick]` -> one JSON line per shape."""
import ctypes as C
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch

from latentsplat_b200 import _capi

CL = torch.channels_last
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
lib = _capi.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

# name, N, Cin, H, W, Cout, k, stride, pad, transposed
SHAPES = [
    ("refine_7x7_128_256@256", 8, 128, 256, 256, 256, 7, 1, 3, 0),
    ("refine_7x7_256_128@256", 8, 256, 256, 256, 128, 7, 1, 3, 0),
    ("convff_7x7_128_256@64", 8, 128, 64, 64, 256, 7, 1, 3, 0),
    ("vae_3x3_512_512@32", 16, 512, 32, 32, 512, 3, 1, 1, 0),
    ("vae_3x3_512_512@64", 16, 512, 64, 64, 512, 3, 1, 1, 0),
    ("vae_3x3_512_512@128", 16, 512, 128, 128, 512, 3, 1, 1, 0),
    ("vae_3x3_256_256@256", 16, 256, 256, 256, 256, 3, 1, 1, 0),
    ("vae_3x3_128_128@256", 16, 128, 256, 256, 128, 3, 1, 1, 0),
    ("vae_out_3x3_128_4@256", 16, 128, 256, 256, 4, 3, 1, 1, 0),
    ("down_4x4s4_128@256", 8, 128, 256, 256, 128, 4, 4, 0, 0),
    ("up_4x4s4_128@64", 8, 128, 64, 64, 128, 4, 4, 0, 1),
    ("gan_4x4s2_64_128@128", 16, 64, 128, 128, 128, 4, 2, 1, 0),
]
if "--quick" in sys.argv:
    SHAPES = SHAPES[:1] + SHAPES[3:5]


def timed(fn, reps=5):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for name, N, Cin, H, W, Cout, k, st, pad, tr in SHAPES:
    x = torch.randn(N, Cin, H, W, device=dev).contiguous(memory_format=CL)
    wshape = (Cin, Cout, k, k) if tr else (Cout, Cin, k, k)
    w = (torch.randn(wshape, device=dev) / (Cin * k * k) ** 0.5).contiguous(memory_format=CL)
    desc = _capi.LsConv2d(N, H, W, Cin, Cout, k, k, st, pad, tr)
    oh, ow = C.c_int32(), C.c_int32()
    _capi.check(lib.ls_conv2d_out_size(C.byref(desc), C.byref(oh), C.byref(ow)), "size")
    y = torch.empty(N, Cout, oh.value, ow.value, device=dev).contiguous(memory_format=CL)
    gy = torch.randn_like(y)
    gx, gw = torch.empty_like(x), torch.empty_like(w)
    s = torch.cuda.current_stream().cuda_stream
    ours = {
        "fwd": timed(lambda: _capi.check(lib.ls_conv2d_forward(C.byref(desc), x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None, 0, s), "f")),
        "dgrad": timed(lambda: _capi.check(lib.ls_conv2d_dgrad(C.byref(desc), gy.data_ptr(), w.data_ptr(), gx.data_ptr(), s), "d")),
        "wgrad": timed(lambda: _capi.check(lib.ls_conv2d_wgrad(C.byref(desc), gy.data_ptr(), x.data_ptr(), gw.data_ptr(), s), "w")),
    }
    conv = lambda: torch.ops.aten.convolution(x, w, None, [st, st], [pad, pad], [1, 1], bool(tr), [0, 0], 1)
    bwd = lambda mask: torch.ops.aten.convolution_backward(gy, x, w, None, [st, st], [pad, pad], [1, 1], bool(tr), [0, 0], 1, mask)
    cudnn = {"fwd": timed(conv), "dgrad": timed(lambda: bwd([True, False, False])), "wgrad": timed(lambda: bwd([False, True, False]))}
    flops = 2.0 * N * oh.value * ow.value * Cout * Cin * k * k if not tr else 2.0 * N * H * W * Cin * Cout * k * k
    print(json.dumps({"shape": name, "gflop": round(flops / 1e9, 1),
                      "ours_ms": {k_: round(v, 4) for k_, v in ours.items()},
                      "ours_tflops": {k_: round(flops / v / 1e9, 1) for k_, v in ours.items()},
                      "cudnn_ms": {k_: round(v, 4) for k_, v in cudnn.items()},
                      "cudnn_tflops": {k_: round(flops / v / 1e9, 1) for k_, v in cudnn.items()}}), flush=True)
