"""One forward (and optionally dgrad / wgrad) launch of one convolution shape, for ncu captures:
   python scripts/conv_one.py N Cin H W Cout k stride pad [fwd|dgrad|wgrad] [reps]"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from latentsplat_b200 import _capi

N, Cin, H, W, Cout, k, st, pad = map(int, sys.argv[1:9])
what = sys.argv[9] if len(sys.argv) > 9 else "fwd"
reps = int(sys.argv[10]) if len(sys.argv) > 10 else 3
dev = torch.device("cuda:0")
CL = torch.channels_last
lib = _capi.load()
x = torch.randn(N, Cin, H, W, device=dev).contiguous(memory_format=CL)
w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).contiguous(memory_format=CL)
b = torch.randn(Cout, device=dev)
desc = _capi.LsConv2d(N, H, W, Cin, Cout, k, k, st, pad, 0)
oh, ow = C.c_int32(), C.c_int32()
_capi.check(lib.ls_conv2d_out_size(C.byref(desc), C.byref(oh), C.byref(ow)), "size")
y = torch.empty(N, Cout, oh.value, ow.value, device=dev).contiguous(memory_format=CL)
gy = torch.randn_like(y)
gx, gw = torch.empty_like(x), torch.empty_like(w)
s = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    if what == "fwd":
        _capi.check(lib.ls_conv2d_forward(C.byref(desc), x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), None, 0, s), "f")
    elif what == "dgrad":
        _capi.check(lib.ls_conv2d_dgrad(C.byref(desc), gy.data_ptr(), w.data_ptr(), gx.data_ptr(), s), "d")
    else:
        _capi.check(lib.ls_conv2d_wgrad(C.byref(desc), gy.data_ptr(), x.data_ptr(), gw.data_ptr(), s), "w")
torch.cuda.synchronize()
print("ok")
