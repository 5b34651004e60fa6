"""GPU probe: raw C-ABI calls of the LSTM kernels in a tight loop (host cost per call ~5 us << kernel time)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_classif_b200 import _lib, ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, T, In, H, L = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (64, 16, 8, 32, 3))]
rnn = torch.nn.LSTM(In, H, num_layers=L, batch_first=True).to(dev)
x = torch.randn(B, T, In, device=dev)
out = torch.empty(L, B, T, H, device=dev); gates = torch.empty(L, B, T, 4 * H, device=dev); cst = torch.empty(L, B, T, H, device=dev)
ps = [[getattr(rnn, f"{n}_l{l}") for l in range(L)] for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
arrs = [ops._ptr_array(p) for p in ps]
st = _lib.stream_ptr()
lib = _lib.lib()
def fwd():
    lib.b2_lstm_stack_fwd(x.data_ptr(), In, arrs[0][1], arrs[1][1], arrs[2][1], arrs[3][1], L, out.data_ptr(), gates.data_ptr(), cst.data_ptr(), B, T, H, st)
dout = torch.randn(B, T, H, device=dev); dx = torch.empty(B, T, In, device=dev)
g = [[torch.zeros_like(p) for p in ps[0]], [torch.zeros_like(p) for p in ps[1]], [torch.zeros(4 * H, device=dev) for _ in range(L)]]
garrs = [ops._ptr_array(t) for t in g]
def bwd():
    lib.b2_lstm_stack_bwd(dout.data_ptr(), x.data_ptr(), In, arrs[0][1], arrs[1][1], L, out.data_ptr(), gates.data_ptr(), cst.data_ptr(), dx.data_ptr(), garrs[0][1], garrs[1][1], garrs[2][1], B, T, H, st)
def timeit(fn, reps=50):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print(f"lstm_stack_fwd  B={B} T={T} In={In} H={H} L={L}: {timeit(fwd):7.1f} us")
print(f"lstm_stack_bwd: {timeit(bwd):7.1f} us")
