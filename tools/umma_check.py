"""developer check of the tcgen05 self-test GEMM against torch (bf16-rounded operands, fp32 accumulate)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autonomous_quadrotor_environment_b200 import _lib as L
lib = L.load_library()
torch.manual_seed(0)
for N, K in [(16, 16), (128, 16), (128, 80), (128, 128), (16, 128), (64, 32)]:
    A = torch.randn(128, K, device="cuda"); B = torch.randn(N, K, device="cuda"); D = torch.zeros(128, N, device="cuda")
    L.check(lib.qs_umma_selftest(N, K, A.data_ptr(), B.data_ptr(), D.data_ptr(), None))
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    err = (D - ref).abs().max().item()
    print("N=%3d K=%3d max|D-ref| = %.3e   (|ref| max %.2f)" % (N, K, err, ref.abs().max().item()), "OK" if err < 1e-3 else "MISMATCH")
