"""Print the phase timeline (SM clocks) of CTA 0's first tile in the tcgen05 layer-forward kernel."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_video_action_b200 import MultiStageModel, _cabi

lib = _cabi.lib()
lens = [4000, 3892, 3600, 3100, 2600, 2000, 1240, 700]
B, T = 8, 4000
net = MultiStageModel(400, 4, 10, 64, 48).cuda()
net.tensor_cores = True
net.pack_ffma_operands = True
with torch.no_grad():
    net(torch.zeros(1, 8, 400, device="cuda"), [8])
x = torch.randn(B * T, 64, device="cuda")
y = torch.empty_like(x); h = torch.empty_like(x)
lens_dev = torch.tensor(lens, dtype=torch.int32, device="cuda")
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
drop = _cabi.MstcnDropout(1, 0, 7, 0)
names = ["start", "setup done", "pdl_wait done", "first TMA issued", "Wd landed (mma)", "full[centre] (mma)",
         "GEMM1 issued+commit", "h_ready seen (mma)", "GEMM2 issued+commit", "full[centre] (epi)", "lo[1] parked",
         "lo[0] parked", "lo[2] parked", "g1_done seen (epi)", "h parked (epi1 done)", "g2_done seen (epi)",
         "tile done (epi2)", "kernel end", "epi2: O loaded", "epi2: y staged", "epi2: rows copied out"]
def off(which, s=1, l=3):
    return C.c_void_p(net._packed.data_ptr() + 4 * lib.mstcn_packed_offset(C.byref(net._dims), s, l, which))
for d in (1,):
    for rep in range(3):
        lib.mstcn_debug_tc_timing(_cabi.ptr(buf))
        _cabi.check(lib.mstcn_layer_fwd_tc(_cabi.ptr(x), _cabi.ptr(y), _cabi.ptr(h), _cabi.ptr(lens_dev), B, T, d,
                                           off(12), off(4), off(6), C.byref(drop), 3, _cabi.stream_ptr()))
        torch.cuda.synchronize()
    lib.mstcn_debug_tc_timing(None)
    t = buf.cpu().tolist()
    print(f"--- dilation {d}")
    for i, n in enumerate(names):
        print(f"{n:28s} {t[i] - t[0]:8d} cyc")
