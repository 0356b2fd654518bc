"""Accuracy of the stem kernels (cg_stem_fwd / cg_stem_wgrad) against torch fp64 on the GPU: the mma.sync path
(split-bf16 operands) and the direct fp32 path (CAUSALGEN_B200_STEM_MMA=0, child process).
Forward: share of bf16 outputs that differ from round_bf16(fp64 result), max |err| in bf16 ulps of the largest output.
Weight gradient: rel-L2 and max rel error vs fp64."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "causal-gen_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def run():
    from causalgen_b200 import _lib as L
    lib = L.load()
    s = torch.cuda.current_stream().cuda_stream
    mode = os.environ.get("CAUSALGEN_B200_STEM_MMA", "1")
    for N, R, cout in ((2, 32, 16), (3, 16, 16), (2, 48, 32), (2, 192, 32), (1, 224, 32)):
        g = torch.Generator().manual_seed(R + cout)
        x8 = torch.randint(0, 256, (N, 1, R, R), generator=g)
        x = ((x8.float() - 127.5) / 127.5).cuda()
        w = (torch.randn(cout, 1, 7, 7, generator=g) * 0.1).cuda()
        b = (torch.randn(cout, generator=g) * 0.1).cuda()
        y = torch.zeros(N, cout // 8, R, R, 8, device="cuda", dtype=torch.bfloat16)
        L.check(lib.cg_stem_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, 1, R, cout, y[0].numel(), s))
        ref = F.conv2d(x.double(), w.double(), b.double(), padding=3)
        got = y.permute(0, 1, 4, 2, 3).reshape(N, cout, R, R).double()
        refb = ref.float().to(torch.bfloat16).double()
        mism = float((got != refb).double().mean())
        err = float((got - ref).abs().max() / ref.abs().max())
        dy = torch.randn(N, cout, R, R, generator=g).cuda().to(torch.bfloat16)
        dyp = dy.reshape(N, cout // 8, 8, R, R).permute(0, 1, 3, 4, 2).contiguous()
        dw, db = torch.zeros_like(w), torch.zeros_like(b)
        L.check(lib.cg_stem_wgrad(x.data_ptr(), dyp.data_ptr(), dw.data_ptr(), db.data_ptr(), N, 1, R, cout, dyp[0].numel(), s))
        wr = w.double().requires_grad_(True)
        br = b.double().requires_grad_(True)
        F.conv2d(x.double(), wr, br, padding=3).backward(dy.double())
        torch.cuda.synchronize()
        print(f"stem_mma={mode} N={N} R={R} cout={cout}: fwd outputs != bf16(fp64) {mism:.5f}, max err {err:.2e} of max |y| "
              f"(bf16 ulp 3.9e-3) | dw rel-L2 {float((dw - wr.grad).norm() / wr.grad.norm()):.2e} "
              f"max {float((dw - wr.grad).abs().max() / wr.grad.abs().max()):.2e} | db rel-L2 "
              f"{float((db - br.grad).norm() / br.grad.norm()):.2e}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        run()
    else:
        for m in ("1", "0"):
            subprocess.check_call([sys.executable, __file__, "--child"], env=dict(os.environ, CAUSALGEN_B200_STEM_MMA=m))
