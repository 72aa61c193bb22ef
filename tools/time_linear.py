"""GPU timing of dmpnn_linear_tc_bf16 vs the SIMT linear (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemprop_b200 import _lib
from chemprop_b200.engine import linear_fwd, linear_tc, pack_weight_tc
R = 502000
for K, lda, N in ((300, 320, 300), (86, 96, 300), (372, 384, 300)):
    A = torch.randn(R, lda, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda") / K ** 0.5
    out = torch.zeros(R, 320, dtype=torch.bfloat16, device="cuda"); Wpk = pack_weight_tc(W)
    def t(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); ts = []
        for _ in range(8):
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]
    tc = t(lambda: linear_tc(A, K, Wpk, N, out, act=_lib.ACT_RELU))
    si = t(lambda: linear_fwd(A, K, W, out, N, act=_lib.ACT_RELU, R=R, pad_to=304))
    fl = 2 * R * K * N; by = R * (K + N) * 2
    print(f"K={K} N={N} R={R}: tc {tc*1e3:.0f} us ({fl/tc/1e9:.0f} TFLOP/s, {by/tc/1e6:.0f} GB/s)  simt {si*1e3:.0f} us ({fl/si/1e9:.0f} TFLOP/s)")
from chemprop_b200.engine import wgrad_tc, linear_wgrad
for K, ldx in ((300, 320), (86, 96), (372, 384)):
    dY = torch.randn(R, 320, device="cuda").bfloat16(); X = torch.randn(R, ldx, device="cuda").bfloat16(); dW = torch.zeros(300, K, device="cuda")
    tc = t(lambda: wgrad_tc(dY, X, R, 300, K, dW)); si = t(lambda: linear_wgrad(dY, X, K, dW, 300, R=R))
    fl = 2 * R * K * 300
    print(f"wgrad K={K}: tc {tc*1e3:.0f} us ({fl/tc/1e9:.0f} TFLOP/s)  simt {si*1e3:.0f} us ({fl/si/1e9:.0f} TFLOP/s)")
