"""A/B of the HBM-bound output-head kernels (out1x1_bwd, head_wgrad, sum_dim0) at c2 sizes: 50 launches between two events, L2-sized
inputs (> 126 MB).  python scripts/head_bench.py [path/to/libgenesis_b200.so]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
path = sys.argv[1] if len(sys.argv) > 1 else 'genesis_b200/lib/libgenesis_b200.so'
lib = ctypes.CDLL(os.path.abspath(path))
P_ = ctypes.c_void_p
def ptr(t): return P_(t.data_ptr())
st = P_(torch.cuda.current_stream().cuda_stream)
N, P, Cin = 320, 4096, 32
NP = N * P
h = torch.randn(NP, Cin, device='cuda'); dh = torch.empty_like(h)
dout = torch.randn(N, 3, P, device='cuda'); out = torch.rand(N, 3, P, device='cuda'); w = torch.randn(3, Cin, device='cuda')
d4 = torch.randn(NP, 4, device='cuda'); dw = torch.zeros(4, Cin, device='cuda')
x = torch.randn(320, 156800, device='cuda'); so = torch.empty(156800, device='cuda')
def timeit(fn, reps=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
lib.g2_out1x1_bwd_f32.argtypes = [P_, P_, P_, P_, P_, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, P_]
lib.g2_head_wgrad_f32.argtypes = [P_, P_, P_, ctypes.c_long, ctypes.c_int, P_]
lib.g2_sum_dim0_f32.argtypes = [P_, P_, ctypes.c_int, ctypes.c_long, P_]
t = timeit(lambda: lib.g2_out1x1_bwd_f32(ptr(dout), ptr(out), ptr(w), ptr(dh), ptr(d4), N, P, Cin, 3, 3, st))
print('%-40s out1x1_bwd  %.4f ms  %6.0f GB/s' % (path[-40:], t, (dh.numel() * 4 + d4.numel() * 4 + 2 * dout.numel() * 4) / t / 1e6))
t = timeit(lambda: lib.g2_head_wgrad_f32(ptr(h), ptr(d4), ptr(dw), NP, Cin, st))
print('%-40s head_wgrad  %.4f ms  %6.0f GB/s' % (path[-40:], t, (h.numel() * 4 + d4.numel() * 4) / t / 1e6))
t = timeit(lambda: lib.g2_sum_dim0_f32(ptr(x), ptr(so), 320, 156800, st))
print('%-40s sum_dim0    %.4f ms  %6.0f GB/s' % (path[-40:], t, x.numel() * 4 / t / 1e6))
