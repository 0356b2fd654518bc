"""CUDA-event timing of the stem kernels at the benchmark shape (128 x 1 x 192 x 192 -> 32 channels); also the ncu target
for them.  usage: python tools/stem_bench.py [N] [R] [reps]   (CAUSALGEN_B200_STEM_MMA=0: the direct fp32 kernels)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
import torch  # noqa: E402
from causalgen_b200 import _lib as L  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
R = int(sys.argv[2]) if len(sys.argv) > 2 else 192
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cout = 32
lib = L.load()
s = torch.cuda.current_stream().cuda_stream
x = torch.rand(N, 1, R, R, device="cuda") * 2 - 1
w = torch.randn(cout, 1, 7, 7, device="cuda") * 0.1
b = torch.randn(cout, device="cuda") * 0.1
y = torch.zeros(N, cout // 8, R, R, 8, device="cuda", dtype=torch.bfloat16)
dy = torch.randn(N, cout // 8, R, R, 8, device="cuda").to(torch.bfloat16)
dw, db = torch.zeros_like(w), torch.zeros_like(b)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2: the timed launches read cold inputs


def timed(fn):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


tf = timed(lambda: L.check(lib.cg_stem_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, 1, R, cout, y[0].numel(), s)))
tw = timed(lambda: L.check(lib.cg_stem_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), N, 1, R, cout, dy[0].numel(), s)))
px = N * R * R
fwd_bytes, wg_bytes = px * (4 + 2 * cout), px * (4 + 2 * cout)
print(f"stem_mma={os.environ.get('CAUSALGEN_B200_STEM_MMA', '1')} N={N} R={R}: fwd {tf:.1f} us ({fwd_bytes / tf / 1e3:.0f} GB/s algorithmic), "
      f"wgrad {tw:.1f} us ({wg_bytes / tw / 1e3:.0f} GB/s)")
